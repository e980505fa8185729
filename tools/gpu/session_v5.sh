set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
( time timeout 2400 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r01_v5.log
timeout 900 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r01_v5.json 2> gpurun_out/bench_r01_v5.log
tail -30 gpurun_out/bench_r01_v5.log
cat gpurun_out/bench_r01_v5.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01_v5_ref.json 2> gpurun_out/bench_r01_v5_ref.log
cat gpurun_out/bench_r01_v5_ref.json
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01_v5.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:trsm_gemm_level -s 30 -c 1 -o gpurun_out/prof_r01_v5_trsm_gemm python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:syrk_mod_kernel -s 3 -c 1 -o gpurun_out/prof_r01_v5_syrk python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out
