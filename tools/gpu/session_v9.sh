set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r01_v9.log
timeout 900 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r01_v9.json 2> gpurun_out/bench_r01_v9.log
tail -62 gpurun_out/bench_r01_v9.log
cat gpurun_out/bench_r01_v9.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01_v9_ref.json 2> gpurun_out/bench_r01_v9_ref.log
cat gpurun_out/bench_r01_v9_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for w in c1 c2; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_r01_v9_$w.json 2> gpurun_out/bench_r01_v9_$w.log; cat gpurun_out/bench_r01_v9_$w.json; done
