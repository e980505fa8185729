# dev: 1536-bit tile kernels compiled for one CTA per SM (255 registers)
set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "1536" 2>&1 | tail -4
( time timeout 900 python bench.py --workload c5-j256-p64-n512-1536b --kernels --steps 2 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_1536_small.json 2> gpurun_out/dev_1536_small.log ) 2>&1 | grep real
grep -E "^\s+\[" gpurun_out/dev_1536_small.log | head -16
( time timeout 900 python bench.py --workload c5-j256-p256-n512-1536b --kernels --steps 2 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_1536_big.json 2> gpurun_out/dev_1536_big.log ) 2>&1 | grep real
grep -E "^\s+\[" gpurun_out/dev_1536_big.log | head -16
python - <<PY
import json
for k in ('small', 'big'):
    d = json.load(open('gpurun_out/dev_1536_%s.json' % k))
    print(k, d['ms_per_step'], d['stages_ms'])
PY
