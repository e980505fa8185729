set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r01_v3.json 2> gpurun_out/bench_r01_v3.log
tail -30 gpurun_out/bench_r01_v3.log
cat gpurun_out/bench_r01_v3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01_v3.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:trsm_gemm_level -s 20 -c 1 -o gpurun_out/prof_r01_v3_trsm_gemm python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out
