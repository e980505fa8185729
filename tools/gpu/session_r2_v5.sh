# round-2 v5 (1 GPU): default bench + reference arm, launch list, DRAM traffic of the trsm stage, ncu --set full of
# the two top kernels, then c4 and the C5 sweep corners
set -x
V=${1:-v5}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
nproc
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
grep -v "^\s*$" gpurun_out/bench_r02_$V.log | head -60
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_${V}_ref.json 2> gpurun_out/bench_r02_${V}_ref.log ) 2>&1 | tail -3
SDPB_B200_CONCURRENCY=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_r02_$V.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"trsm_|syrk_" -c 120 --csv --log-file gpurun_out/traffic_r02_$V.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"trsm_gemm_level2" -s 14 -c 7 -o /tmp/trsm_full python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
ncu -i /tmp/trsm_full.ncu-rep --page raw --csv > gpurun_out/prof_r02_${V}_trsm_gemm_raw.csv
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"syrk_imma" -s 3 -c 1 -o /tmp/syrk_full python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
ncu -i /tmp/syrk_full.ncu-rep --page raw --csv > gpurun_out/prof_r02_${V}_syrk_imma_raw.csv
for w in c1 c2 c4 c5-j256-p64-n512 c5-j256-p64-n512-1536b c5-j256-p256-n512 c5-j1024-p256-n512 c5-j256-p64-n4096-256b c5-j256-p64-n4096 c5-j256-p256-n512-1536b c5-j256-p64-n4096-1536b; do
  ( time timeout 900 python bench.py --workload $w --kernels --steps 2 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_${V}_$w.json 2> gpurun_out/bench_r02_${V}_$w.log ) 2>&1 | grep real
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_${V}_$w.json'))
    print('$w', 'ms/step', round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['value'] * 1e3, 1), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'], 4),
          'int', round((d['roofline'].get('int_pipe') or {}).get('frac', 0), 3), 'solve', round(d['schur_solve']['device_ms'], 1), d['stages_ms'])
except Exception as e:
    print('$w', 'FAILED', e)
PY
done
ls -la gpurun_out | tail -5
