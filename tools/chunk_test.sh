cd /root/repo
for cfg in "default::" "k5geom:5:1" "k4geom:4:1" "k6geom:6:1" "k8eq:8:0"; do
  name=${cfg%%:*}; rest=${cfg#*:}; k=${rest%%:*}; g=${rest#*:}
  if [ -n "$k" ]; then export BA_PIPELINE_CHUNKS=$k BA_PIPELINE_GEOM=$g; else unset BA_PIPELINE_CHUNKS BA_PIPELINE_GEOM; fi
  BA_BENCH_NO_STRONG=1 python bench.py --pairs 25000 --steps 4 --warmup 2 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$name', 'kernel %.2f ms  e2e %.2f ms (%.1f GCUPS)'%(l['ms_per_step'], l['e2e']['ms_per_step'], l['e2e']['value']))"
done
