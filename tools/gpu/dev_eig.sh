# dev: step_length on the device (row N3) -- parity tests at the precisions of the development build, bench
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "(search_direction_bit_exact and (768 or 664)) or step_length" 2>&1 | tail -25
timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_eig_bench.json 2> gpurun_out/dev_eig_bench.log
python - <<PY
import json
d = json.load(open('gpurun_out/dev_eig_bench.json'))
print(d['ms_per_step'], d['step_length'])
print(d['search_direction']['device_ms'])
PY
