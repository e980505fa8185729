# dev: exact syrk on the integer tensor path -- parity at the precisions of the development build, then bench A/B
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded_local.py -x -q -k "768 or 664 or c3 or sharded or local" 2>&1 | tail -15
timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_syrk_imma.json 2> gpurun_out/dev_syrk_imma.log
grep -E "syrk|stages" gpurun_out/dev_syrk_imma.log
SDPB_B200_SYRK=imad timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_syrk_imad.json 2> gpurun_out/dev_syrk_imad.log
grep -E "syrk|stages" gpurun_out/dev_syrk_imad.log
python - <<PY
import json
for k in ('imma','imad'):
    d = json.load(open('gpurun_out/dev_syrk_%s.json' % k))
    print(k, d['ms_per_step'], d['stages_ms'])
PY
