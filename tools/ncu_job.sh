#!/bin/bash
# ncu evidence for one workload: launch list (gpu__time_duration) + one --set full capture of the alignment kernel.
#   tools/ncu_job.sh <workload> <pairs> <tag>    -> gpurun_out/ncu_<tag>_launches.csv, gpurun_out/ncu_<tag>.ncu-rep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
w=$1; pairs=$2; tag=$3
export BA_BENCH_NO_STRONG=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ncu_${tag}_launches.csv \
  python bench.py --workload $w --pairs $pairs --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_${tag}_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ba_align_kernel -s 1 -c 1 -f -o gpurun_out/ncu_${tag} \
  python bench.py --workload $w --pairs $pairs --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_${tag}_full.log 2>&1
ls -la gpurun_out/ncu_${tag}*
