#!/bin/bash
# Batch-traceback launch shapes on C5: tools/variant_tb.sh <variant>:<BA_TB_STRIDE or -> ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  v=${spec%%:*}; st=${spec#*:}
  if [ "$st" = "-" ]; then unset BA_TB_STRIDE; else export BA_TB_STRIDE=$st; fi
  BLOCK_ALIGNER_B200_LIB=$PWD/build/variants/libba_$v.so BA_BENCH_NO_STRONG=1 timeout 600 python bench.py --steps 2 --warmup 2 --workload C5_longread_trace_50k \
    --no-cpu-baseline --no-configs > gpurun_out/tbv_${v}_$st.json 2> gpurun_out/tbv_${v}_$st.err
  python - "$v" "$st" <<'PY'
import json, sys
v, st = sys.argv[1], sys.argv[2]
try:
    l = json.loads(open(f"gpurun_out/tbv_{v}_{st}.json").read().strip().split("\n")[-1])
    print(f"{v} stride {st}: kernel {l['value']:.1f} GCUPS ({l['ms_per_step']:.2f} ms)  e2e {l['e2e']['value']:.1f} failed {l['n_failed_pairs']}")
except Exception as e:
    print(v, st, "FAILED", e)
PY
done
