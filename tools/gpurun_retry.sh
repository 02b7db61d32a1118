#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient"): tools/gpurun_retry.sh <timeout> '<command>'
t=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|no box\|busy"; then sleep 120; continue; fi
  break
done
