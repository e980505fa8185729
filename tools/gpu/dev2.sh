for g in 1 2 4; do
  echo "GROUPS=$g"
  SDPB_B200_GROUPS=$g timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['serial_ms_per_step'], d['stages_ms'])
"
done
