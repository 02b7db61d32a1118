#!/bin/bash
# C5 (TRACE) numbers of tuning variants: tools/variant_c5.sh <name> [<name> ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  BLOCK_ALIGNER_B200_LIB=$PWD/build/variants/libba_$v.so BA_BENCH_NO_STRONG=1 timeout 600 python bench.py --steps 2 --warmup 2 --workload C5_longread_trace_50k \
    --cpu-seconds 2 --no-configs > gpurun_out/variant_${v}_C5.json 2> gpurun_out/variant_${v}_C5.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    l = json.loads(open(f"gpurun_out/variant_{v}_C5.json").read().strip().split("\n")[-1])
    print(f"{v}_C5: kernel {l['value']:.1f} GCUPS ({l['ms_per_step']:.2f} ms)  e2e {l['e2e']['value']:.1f}  failed {l['n_failed_pairs']} parity {l.get('parity',{}).get('mismatches')}/{l.get('parity',{}).get('pairs_checked')}")
except Exception as e:
    print(v, "FAILED", e)
PY
done
