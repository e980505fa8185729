# 8 GPUs of one box: c3 weak scaling (e2e with resident factors), then c4 with the panel-distributed Cholesky(Q)
set -x
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv | head -10
nproc; free -g | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_r02_c3_n$N.json 2> gpurun_out/scale_r02_c3_n$N.log
grep -v "^\s*$" gpurun_out/scale_r02_c3_n$N.log | tail -5
python - <<PY
import json
d = json.load(open('gpurun_out/scale_r02_c3_n$N.json'))
print('c3 N=$N ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'all', d['e2e_all_outputs']['value'], 'solve', d['schur_solve']['device_ms'], d['schur_solve']['kernels_ms'], 'numa', d['numa'], 'stages', d['stages_ms'])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --workload c4 --gpus $N --steps 3 --warmup 3 --no-cpu --no-all-outputs --kernels > gpurun_out/scale_r02_c4_n$N.json 2> gpurun_out/scale_r02_c4_n$N.log
grep -E "^\s+\[|stages|nccl|rror" gpurun_out/scale_r02_c4_n$N.log | head -40
python - <<PY
import json
d = json.load(open('gpurun_out/scale_r02_c4_n$N.json'))
print('c4 N=$N ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'solve', d['schur_solve']['device_ms'], 'stages', d['stages_ms'])
PY
