# dev: Cholesky pivots with the reciprocal seeded from the square root's rsqrt iterate
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "(cooperative_pivot and 768) or (schur_step_bit_exact and (768 or 664)) or c3_sample" 2>&1 | tail -8
timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_pivot.json 2> gpurun_out/dev_pivot.log
grep -E "potrf|stages" gpurun_out/dev_pivot.log
python - <<PY
import json
d = json.load(open('gpurun_out/dev_pivot.json'))
print(d['ms_per_step'], d['stages_ms'])
PY
