# dev: S chain split by size class (default for two-class batches) against one group
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded_local.py -x -q -k "block_groups or c3 or named_config or sharded or local" 2>&1 | tail -6
timeout 600 python bench.py --kernels --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_groups_size.json 2> gpurun_out/dev_groups_size.log
SDPB_B200_GROUPS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_groups_one.json 2> /dev/null
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_groups_size_c4.json 2> /dev/null
python - <<PY
import json
for k in ('size', 'one', 'size_c4'):
    d = json.load(open('gpurun_out/dev_groups_%s.json' % k))
    print(k, d['ms_per_step'], d['e2e']['value'], d.get('serial_ms_per_step'), d['stages_ms'])
PY
