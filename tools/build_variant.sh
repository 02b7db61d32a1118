#!/bin/bash
# Tuning build: only the C2 kernels, extra nvcc flags from the command line.
#   tools/build_variant.sh <name> [-DBA_LB_BLOCKS=5 ...]   ->  build/variants/libba_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC -Xcompiler -O3 -shared -DBA_MINIMAL "$@" -o build/variants/libba_$name.so \
  block_aligner_b200/csrc/ba_runtime.cu block_aligner_b200/csrc/ba_launch.cu block_aligner_b200/csrc/ba_capi.cpp -lcudart
