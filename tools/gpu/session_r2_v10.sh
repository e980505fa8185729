# round-2 v10 (1 GPU): size split on urgent streams by default, block-diagonal solves on the tile
# kernels, out-of-line wmul/wadd in the pivot -- all GPU tests, the c3 bench, A/B of the block-diagonal
# solves, c1 / c2 / c4, and the source-level capture of the Cholesky(Q) diagonal kernel again
set -x
V=${1:-v10}
( time timeout 1500 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_r02_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
grep -v "^\s*$" gpurun_out/bench_r02_$V.log | head -48
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_$V.json'))
print('$V', d['ms_per_step'], d['e2e'], d['stages_ms'], d['search_direction']['device_ms'], d['step_length']['device_ms'], d['step_length'].get('kernels_ms_primal'), d['e2e_newton_iteration']['value'])
print(d['search_direction'].get('kernels_ms_predictor'))
PY
SDPB_B200_BDM_TRSM=rl timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_bdm_rl.json 2> gpurun_out/dev_bdm_rl.log
python - <<PY
import json
d = json.load(open('gpurun_out/dev_bdm_rl.json'))
print('bdm rl', d['ms_per_step'], d['search_direction']['device_ms'], d['step_length']['device_ms'], d['e2e_newton_iteration']['value'])
PY
for w in c1 c2 c4; do
  timeout 600 python bench.py --workload $w --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_${V}_$w.json 2> gpurun_out/bench_r02_${V}_$w.log
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_${V}_$w.json'))
print('$w', 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'] * 1e3, 2), d['stages_ms'], 'newton', d['e2e_newton_iteration']['value'])
PY
done
SDPB_B200_GROUPS=1 timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_c4_g1.json 2> gpurun_out/dev_c4_g1.log
python -c "import json; d=json.load(open('gpurun_out/dev_c4_g1.json')); print('c4 one group', d['ms_per_step'], d['e2e']['value'])"
SDPB_B200_CONCURRENCY=0 timeout 600 ncu --set full --import-source on --warp-sampling-interval 1 --clock-control none -k regex:potrf_diag_rl -s 25 -c 1 -o /tmp/diag_rl python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_ncu.log 2>&1
ncu -i /tmp/diag_rl.ncu-rep --page source --print-source sass --csv > gpurun_out/prof_r02_${V}_potrf_diag_rl_source.csv
ncu -i /tmp/diag_rl.ncu-rep --page raw --csv > gpurun_out/prof_r02_${V}_potrf_diag_rl_raw.csv
ls -la gpurun_out | tail -4
