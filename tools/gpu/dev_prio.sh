# dev: chol(X) / chol(Y) on streams of the greatest priority (A/B inside one session), the
# Cholesky(Q) diagonal kernel compiled for one CTA per SM, and a source-level capture of it
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "schur_step_bit_exact or c3_sample or separate_calls" 2>&1 | tail -4
for p in 0 1 0 1; do
  SDPB_B200_PRIORITY=$p timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_prio$p.json 2> gpurun_out/dev_prio$p.log
  python - <<PY
import json
d = json.load(open('gpurun_out/dev_prio$p.json'))
print('priority $p', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'] * 1e3, 2), 'serial', d.get('serial_ms_per_step'))
PY
done
for p in 0 1; do
  SDPB_B200_GROUPS=size SDPB_B200_PRIORITY=$p timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_prio_size$p.json 2> gpurun_out/dev_prio_size$p.log
  python - <<PY
import json
d = json.load(open('gpurun_out/dev_prio_size$p.json'))
print('size split, priority $p', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'] * 1e3, 2))
PY
done
SDPB_B200_GROUPS=size timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "schur_step_bit_exact or c3_sample" 2>&1 | tail -3
for w in c1 c2; do for tb in 0 32768; do
  SDPB_B200_TRSM_TILE_BELOW=$tb timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_tile_${w}_$tb.json 2> gpurun_out/dev_tile_${w}_$tb.log
  python - <<PY
import json
d = json.load(open('gpurun_out/dev_tile_${w}_$tb.json'))
print('$w tile_below $tb', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'] * 1e3, 3), d['stages_ms'])
PY
done; done
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden_trajectory.py -x -q -m gpu -k "named_config or golden or small or ragged" 2>&1 | tail -3
timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_prio_k.json 2> gpurun_out/dev_prio_k.log
grep -E "potrf_Q|stages" gpurun_out/dev_prio_k.log
SDPB_B200_CONCURRENCY=0 timeout 600 ncu --set full --import-source on --warp-sampling-interval 1 --clock-control none -k regex:potrf_diag_rl -s 25 -c 1 -o /tmp/diag_rl python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_ncu.log 2>&1
tail -5 gpurun_out/dev_ncu.log
ncu -i /tmp/diag_rl.ncu-rep --page source --print-source sass --csv > gpurun_out/prof_r02_v9_potrf_diag_rl_source.csv
ncu -i /tmp/diag_rl.ncu-rep --page raw --csv > gpurun_out/prof_r02_v9_potrf_diag_rl_raw.csv
ls -la gpurun_out | tail -4
