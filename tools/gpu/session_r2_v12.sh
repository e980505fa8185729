# round-2 v12 (1 GPU): Cholesky(Q) with the pipelined diagonal + panel kernel -- the new tests first
# (under a short timeout), then all GPU tests, smoke, the c3 bench, A/B against the two-kernel form, c1 / c2 / c4
set -x
V=${1:-v12}
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "cholesky_Q_pipelined or rank_deficient or c3_sample" 2>&1 | tail -5
( time timeout 1200 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
grep -E "potrf_Q|stages" gpurun_out/bench_r02_$V.log
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_$V.json'))
print('$V', d['ms_per_step'], d['e2e'], d['stages_ms'], d['e2e_newton_iteration']['value'], d['cpu_baseline']['value'])
PY
SDPB_B200_POTRF_FUSED=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_unfused.json 2> gpurun_out/dev_unfused.log
python -c "import json; d=json.load(open('gpurun_out/dev_unfused.json')); print('two kernels', d['ms_per_step'], d['e2e']['value'])"
for w in c1 c2 c4; do
  timeout 600 python bench.py --workload $w --kernels --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_${V}_$w.json 2> gpurun_out/bench_r02_${V}_$w.log
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_${V}_$w.json'))
print('$w', 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'] * 1e3, 2), d['stages_ms'])
PY
done
SDPB_B200_CONCURRENCY=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_r02_$V.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
ls -la gpurun_out | tail -3
