#!/bin/bash
# Batch-traceback tuning: C5 kernel time for several numbers of walks in flight (BA_TB_WALKS) and the unstaged kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for wk in "$@"; do
  if [ "$wk" = old ]; then export BA_TB_STRIDE=8; unset BA_TB_WALKS; else unset BA_TB_STRIDE; export BA_TB_WALKS=$wk; fi
  BA_BENCH_NO_STRONG=1 timeout 600 python bench.py --workload C5_longread_trace_50k --steps 2 --warmup 2 --cpu-seconds 2 --no-configs > gpurun_out/tb_$wk.json 2> gpurun_out/tb_$wk.err
  python - "$wk" <<'PY'
import json, sys
wk = sys.argv[1]
try:
    l = json.loads(open(f"gpurun_out/tb_{wk}.json").read().strip().split("\n")[-1])
    print(f"walks {wk}: kernel {l['value']:.1f} GCUPS ({l['ms_per_step']:.2f} ms)  e2e {l['e2e']['value']:.1f}  parity {l.get('parity',{}).get('mismatches')}/{l.get('parity',{}).get('pairs_checked')}")
except Exception as e:
    print(wk, "FAILED", e)
PY
done
