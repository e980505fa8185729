set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r01_v8.log
timeout 900 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r01_v8.json 2> gpurun_out/bench_r01_v8.log
tail -60 gpurun_out/bench_r01_v8.log
cat gpurun_out/bench_r01_v8.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
