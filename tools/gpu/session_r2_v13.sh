# round-2 v13 (1 GPU, the last session of the round): Barrett reduction in the Garner recurrence of
# crt_restore_kernel -- all GPU tests, smoke, c3 / c1 / c2 bench lines
set -x
V=${1:-v13}
( time timeout 600 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --kernels --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
grep -E "crt_restore|stages" gpurun_out/bench_r02_$V.log
python -c "import json; d=json.load(open('gpurun_out/bench_r02_$V.json')); print('$V', d['ms_per_step'], d['e2e'], d['e2e_newton_iteration']['value'])"
for w in c1 c2; do
  timeout 200 python bench.py --workload $w --kernels --steps 10 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_${V}_$w.json 2> gpurun_out/bench_r02_${V}_$w.log
  python -c "import json; d=json.load(open('gpurun_out/bench_r02_${V}_$w.json')); print('$w', round(d['ms_per_step'],3), round(d['e2e']['value']*1e3,3), d['stages_ms'])"
done
