#!/bin/bash
# One gpurun call of round 2: GPU tests, default bench (all five configs), launch list, GPU soak.
# usage: tools/gpu_job.sh [soak minutes]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_box.txt 2>&1
nproc >> gpurun_out/gpu_box.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
(time timeout 900 python bench.py) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_bench.log 2>&1
SOAK=${1:-3}
timeout $((SOAK * 60 + 120)) python tools/soak.py cuda $SOAK 1 > gpurun_out/soak_gpu.txt 2>&1
tail -3 gpurun_out/soak_gpu.txt
