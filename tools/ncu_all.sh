#!/bin/bash
# ncu evidence for every BASELINE workload at full size: launch list + one --set full capture of the alignment kernel
# (and of the batch traceback kernel for C5). Results under gpurun_out/.
cd "$(dirname "$0")/.."
bash tools/ncu_job.sh C2_nanopore_xdrop_10k 100000 r02_C2
bash tools/ncu_job.sh C3_uniclust_protein_global 1000000 r02_C3
bash tools/ncu_job.sh C4_seq_to_profile_xdrop 100000 r02_C4
bash tools/ncu_job.sh C5_longread_trace_50k 10000 r02_C5
BA_BENCH_NO_STRONG=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ba_traceback_batch -s 1 -c 1 -f -o gpurun_out/ncu_r02_C5_traceback \
  python bench.py --workload C5_longread_trace_50k --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_r02_C5_traceback.log 2>&1
ls -la gpurun_out/*.ncu-rep
