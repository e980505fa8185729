set -x
( time timeout 2400 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r01_v6.log
timeout 900 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r01_v6.json 2> gpurun_out/bench_r01_v6.log
tail -22 gpurun_out/bench_r01_v6.log
cat gpurun_out/bench_r01_v6.json
for w in c1 c2 c4; do
  timeout 1200 python bench.py --workload $w --kernels --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r01_v6_$w.json 2> gpurun_out/bench_r01_v6_$w.log
  tail -22 gpurun_out/bench_r01_v6_$w.log | cut -c1-120
  cat gpurun_out/bench_r01_v6_$w.json | cut -c1-400
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
