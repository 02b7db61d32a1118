#!/bin/bash
# One workload on tuning variants: tools/variant_wl.sh <workload> <name> [<name> ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WL=$1; shift
for v in "$@"; do
  BLOCK_ALIGNER_B200_LIB=$PWD/build/variants/libba_$v.so BA_BENCH_NO_STRONG=1 timeout 600 python bench.py --steps 2 --warmup 2 --workload $WL \
    --cpu-seconds 2 --no-configs > gpurun_out/variant_${v}_$WL.json 2> gpurun_out/variant_${v}_$WL.err
  python - "$v" "$WL" <<'PY'
import json, sys
v, wl = sys.argv[1], sys.argv[2]
try:
    l = json.loads(open(f"gpurun_out/variant_{v}_{wl}.json").read().strip().split("\n")[-1])
    print(f"{v} {wl}: kernel {l['value']:.1f} GCUPS ({l['ms_per_step']:.2f} ms)  e2e {l['e2e']['value']:.1f}  failed {l['n_failed_pairs']} parity {l.get('parity',{}).get('mismatches')}/{l.get('parity',{}).get('pairs_checked')}")
except Exception as e:
    print(v, "FAILED", e)
PY
done
