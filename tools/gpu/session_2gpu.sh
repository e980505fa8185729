set -x
( time timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r01_v10_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu > gpurun_out/scale_r01_v10_n2.json 2> gpurun_out/scale_r01_v10_n2.log
grep -v "^\s*$" gpurun_out/scale_r01_v10_n2.log | tail -12
cat gpurun_out/scale_r01_v10_n2.json
