set -x
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --workload c4 --gpus $N --kernels --steps 2 --warmup 3 --no-cpu > gpurun_out/scale_r01_v9_c4_n$N.json 2> gpurun_out/scale_r01_v9_c4_n$N.log
grep -v "^\s*$" gpurun_out/scale_r01_v9_c4_n$N.log | grep -i "potrf_Q\|nccl\|panel_pack\|stages\|solve\|setup\|error\|Traceback" | head -40
cat gpurun_out/scale_r01_v9_c4_n$N.json
