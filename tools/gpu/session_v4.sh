set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r01_v4.log
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r01_v4.json 2> gpurun_out/bench_r01_v4.log
tail -40 gpurun_out/bench_r01_v4.log
cat gpurun_out/bench_r01_v4.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01_v4_ref.json 2> gpurun_out/bench_r01_v4_ref.log
cat gpurun_out/bench_r01_v4_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01_v4.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out
