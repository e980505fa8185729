# dev: k-chunk of 8 rows between CTA barriers in the tile kernels
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "(schur_step_bit_exact and (768 or 664)) or c3_sample or (search_direction and 768)" 2>&1 | tail -5
timeout 600 python bench.py --kernels --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_kc8.json 2> gpurun_out/dev_kc8.log
grep -E "^\s+\[" gpurun_out/dev_kc8.log | head -12
python - <<PY
import json
d = json.load(open('gpurun_out/dev_kc8.json'))
print(d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d['roofline']['int_pipe']['frac'])
PY
