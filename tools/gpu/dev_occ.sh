# dev: c4 (960 bits, NL = 17) with the tile kernels at one CTA per SM; c3 with the re-ordered IMMA issue
set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "960 or c4 or c3_sample" 2>&1 | tail -4
( time timeout 900 python bench.py --workload c4 --kernels --steps 2 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_occ_c4.json 2> gpurun_out/dev_occ_c4.log ) 2>&1 | grep real
grep -E "^\s+\[" gpurun_out/dev_occ_c4.log | head -12
timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_occ_c3.json 2> gpurun_out/dev_occ_c3.log
grep -E "syrk|stages" gpurun_out/dev_occ_c3.log
python - <<PY
import json
for k in ('c4', 'c3'):
    d = json.load(open('gpurun_out/dev_occ_%s.json' % k))
    print(k, d['ms_per_step'], d['stages_ms'])
PY
