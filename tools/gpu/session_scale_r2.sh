# round-2 multi-GPU session: N GPUs of one box.  NCCL parity tests (N >= 2), c3 weak scaling, c4 at N = 8.
set -x
N=${1:-2}
V=${2:-v5}
nvidia-smi --query-gpu=index,name --format=csv | head -10
nproc
if [ "$N" = "2" ]; then
  ( time timeout 1200 python -m pytest tests/test_sharded.py -m gpu -q ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r02_${V}_2gpu.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_r02_${V}_c3_n$N.json 2> gpurun_out/scale_r02_${V}_c3_n$N.log
python - <<PY
import json
d = json.load(open('gpurun_out/scale_r02_${V}_c3_n$N.json'))
print('c3 N=$N ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'all', (d.get('e2e_all_outputs') or {}).get('value'), 'solve', d['schur_solve']['device_ms'], 'stages', d['stages_ms'])
print('direction', d['search_direction'].get('api_ms_host_buffers'), 'step_length', d['step_length'].get('api_ms_both_calls'))
PY
if [ "$N" = "8" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --workload c4 --gpus $N --steps 3 --warmup 3 --no-cpu --no-all-outputs --kernels > gpurun_out/scale_r02_${V}_c4_n$N.json 2> gpurun_out/scale_r02_${V}_c4_n$N.log
  grep -a -E "^\s+\[|nccl" gpurun_out/scale_r02_${V}_c4_n$N.log | head -30
  python - <<PY
import json
d = json.load(open('gpurun_out/scale_r02_${V}_c4_n$N.json'))
print('c4 N=$N ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'solve', d['schur_solve']['device_ms'], 'stages', d['stages_ms'])
PY
fi
