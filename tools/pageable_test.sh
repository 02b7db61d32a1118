#!/bin/bash
# End-to-end C2 with caller buffers in pageable memory for several numbers of staging threads (BA_STAGE_THREADS)
cd "$(dirname "$0")/.."
for t in "$@"; do
  BA_STAGE_THREADS=$t BA_BENCH_NO_STRONG=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().split('\n')[-1]); p=l['e2e'].get('pageable_caller_buffers',{}); print('threads $t: pinned e2e %.1f ms, pageable %.1f GCUPS (%.1f ms)'%(l['e2e']['ms_per_step'], p.get('value',0), p.get('ms_per_step',0)))"
done
