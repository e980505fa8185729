# dev: lazily allocated direction / scale_multiply_add regions; the largest C5 corner
set -x
true
( time timeout 900 python bench.py --workload c5-j1024-p256-n512 --kernels --steps 2 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_v4_c5-j1024-p256-n512.json 2> gpurun_out/bench_r02_v4_c5-j1024-p256-n512.log ) 2>&1 | grep real
tail -3 gpurun_out/bench_r02_v4_c5-j1024-p256-n512.log | cut -c1-300
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_v4_c5-j1024-p256-n512.json'))
print(d['ms_per_step'], d['e2e'], d['stages_ms'], d.get('search_direction', {}).get('api_ms_host_buffers'))
PY
