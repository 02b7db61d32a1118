#!/bin/bash
# Kernel + end-to-end numbers of the in-tree library for a list of workloads (parity against the oracle included).
#   tools/quick_bench.sh <tag> <workload> [<workload> ...]   -> gpurun_out/quick_<tag>_<workload>.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
for w in "$@"; do
  BA_BENCH_NO_STRONG=1 timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --cpu-seconds 3 --no-configs > gpurun_out/quick_${tag}_$w.json 2> gpurun_out/quick_${tag}_$w.err
  python - "$tag" "$w" <<'PY'
import json, sys
tag, w = sys.argv[1], sys.argv[2]
try:
    l = json.loads(open(f"gpurun_out/quick_{tag}_{w}.json").read().strip().split("\n")[-1])
    print(f"{w}: kernel {l['value']:.1f} GCUPS ({l['ms_per_step']:.2f} ms)  e2e {l['e2e']['value']:.1f}  failed {l['n_failed_pairs']} parity {l.get('parity',{}).get('mismatches')}/{l.get('parity',{}).get('pairs_checked')}")
except Exception as e:
    print(w, "FAILED", e)
PY
done
