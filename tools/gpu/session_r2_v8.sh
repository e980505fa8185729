# round-2 v8 (1 GPU): division decided by the guard word (mpfw::div_recip) -- all GPU tests, the
# c3 bench with the per-kernel pass, c1/c2, and one source-level capture of the Cholesky(Q) diagonal
# kernel (where the warp-cooperative pivot spends its time)
set -x
V=${1:-v8}
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
grep -v "^\s*$" gpurun_out/bench_r02_$V.log | head -45
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_$V.json'))
print('$V', d['ms_per_step'], d['e2e'], d['stages_ms'], d['search_direction']['device_ms'], d['step_length']['device_ms'], d['e2e_newton_iteration']['value'])
PY
for w in c1 c2; do
  timeout 600 python bench.py --workload $w --kernels --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_${V}_$w.json 2> gpurun_out/bench_r02_${V}_$w.log
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_${V}_$w.json'))
print('$w', 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'] * 1e3, 2), d['stages_ms'])
PY
done
SDPB_B200_CONCURRENCY=0 timeout 600 ncu --set full --import-source on --warp-sampling-interval 1 --clock-control none -k regex:potrf_diag_rl -s 25 -c 1 -o /tmp/diag_rl python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
ncu -i /tmp/diag_rl.ncu-rep --page source --print-source sass --csv > gpurun_out/prof_r02_${V}_potrf_diag_rl_source.csv
ncu -i /tmp/diag_rl.ncu-rep --page source --print-source cuda --csv > gpurun_out/prof_r02_${V}_potrf_diag_rl_cuda.csv 2>&1
ncu -i /tmp/diag_rl.ncu-rep --page raw --csv > gpurun_out/prof_r02_${V}_potrf_diag_rl_raw.csv
ls -la gpurun_out | tail -5
