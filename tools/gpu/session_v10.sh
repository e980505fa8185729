set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r01_v10.log
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r01_v10.json 2> gpurun_out/bench_r01_v10.log
tail -12 gpurun_out/bench_r01_v10.log
cat gpurun_out/bench_r01_v10.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01_v10_ref.json 2> gpurun_out/bench_r01_v10_ref.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
