# round-2 v3: full GPU test suite, smoke, default bench + reference arm, launch list and DRAM traffic of the trsm stage
set -x
V=${1:-v3}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
nproc
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r02_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
grep -v "^\s*$" gpurun_out/bench_r02_$V.log | head -45
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_${V}_ref.json 2> gpurun_out/bench_r02_${V}_ref.log ) 2>&1 | tail -3
SDPB_B200_CONCURRENCY=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r02_$V.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"trsm_" -c 200 --csv --log-file gpurun_out/traffic_trsm_r02_$V.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
ls -la gpurun_out | tail -8
