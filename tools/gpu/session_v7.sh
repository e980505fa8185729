set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
( time timeout 2400 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r01_v7.log
timeout 900 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r01_v7.json 2> gpurun_out/bench_r01_v7.log
tail -45 gpurun_out/bench_r01_v7.log
cat gpurun_out/bench_r01_v7.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01_v7_ref.json 2> gpurun_out/bench_r01_v7_ref.log
cat gpurun_out/bench_r01_v7_ref.json
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01_v7.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --set full --clock-control none -k regex:"trsm_(gemm|diag)_level" -s 65 -c 15 --csv --page raw --log-file gpurun_out/traffic_trsm_r01_v7.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:trsm_gemm_level -s 30 -c 1 -o gpurun_out/prof_r01_v7_trsm_gemm python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ls -la gpurun_out | tail -12
