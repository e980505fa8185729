# dev: right-looking block triangular solves in the direction and in step_length
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded_local.py -x -q -k "(search_direction and (768 or 664)) or step_length or sharded_matches" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_golden_trajectory.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_rl.json 2> gpurun_out/dev_rl.log
python - <<PY
import json
d = json.load(open('gpurun_out/dev_rl.json'))
print(d['ms_per_step'])
print(d['search_direction'])
print(d['step_length'])
PY
