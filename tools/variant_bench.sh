#!/bin/bash
# Bench tuning variants built by tools/build_variant.sh (C2 / C3 / C5 kernels only) on the GPU box.
#   tools/variant_bench.sh <pairs> <name> [<name> ...]   -> gpurun_out/variant_<name>.json (one bench line each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
pairs=$1; shift
for v in "$@"; do
  BLOCK_ALIGNER_B200_LIB=$PWD/build/variants/libba_$v.so BA_BENCH_NO_STRONG=1 timeout 600 python bench.py --steps 3 --warmup 2 \
    --pairs $pairs --cpu-seconds 3 --no-configs > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  BLOCK_ALIGNER_B200_LIB=$PWD/build/variants/libba_$v.so BA_BENCH_NO_STRONG=1 timeout 600 python bench.py --steps 3 --warmup 2 --workload C3_uniclust_protein_global \
    --pairs 400000 --cpu-seconds 2 --no-configs > gpurun_out/variant_${v}_C3.json 2> gpurun_out/variant_${v}_C3.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
for suffix in ("", "_C3"):
    try:
        l = json.loads(open(f"gpurun_out/variant_{v}{suffix}.json").read().strip().split("\n")[-1])
        print(f"{v}{suffix}: kernel {l['value']:.1f} GCUPS ({l['ms_per_step']:.2f} ms)  e2e {l['e2e']['value']:.1f}  failed {l['n_failed_pairs']} parity {l.get('parity',{}).get('mismatches')}/{l.get('parity',{}).get('pairs_checked')}")
    except Exception as e:
        print(v + suffix, "FAILED", e)
PY
done
