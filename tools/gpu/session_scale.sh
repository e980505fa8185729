set -x
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu > gpurun_out/scale_r01_v10_n$N.json 2> gpurun_out/scale_r01_v10_n$N.log
grep -v "^\s*$" gpurun_out/scale_r01_v10_n$N.log | tail -12
cat gpurun_out/scale_r01_v10_n$N.json
