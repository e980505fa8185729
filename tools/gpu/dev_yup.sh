# dev: Y uploaded on the copy stream under the X chain
set -x
( time timeout 1500 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/dev_yup.json 2> gpurun_out/dev_yup.log
python - <<PY
import json
d = json.load(open('gpurun_out/dev_yup.json'))
print(d['ms_per_step'], d['e2e'], d['e2e_all_outputs']['value'])
PY
