# round-2: full GPU test suite, default bench + reference arm, then c4 and the C5 sweep corners on one GPU
set -x
V=${1:-v2}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
nproc; free -g | head -2
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r02_$V.log
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
grep -v "^\s*$" gpurun_out/bench_r02_$V.log | head -45
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_${V}_ref.json 2> gpurun_out/bench_r02_${V}_ref.log ) 2>&1 | tail -3
tail -3 gpurun_out/bench_r02_${V}_ref.log
for w in c4 c5-j256-p64-n512 c5-j256-p256-n512 c5-j1024-p256-n512 c5-j256-p64-n4096-256b c5-j256-p64-n4096 c5-j256-p256-n512-1536b c5-j256-p64-n4096-1536b; do
  ( time timeout 900 python bench.py --workload $w --kernels --steps 2 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_${V}_$w.json 2> gpurun_out/bench_r02_${V}_$w.log ) 2>&1 | grep real
  grep -E "^\s+\[|stages|Error|error|rror" gpurun_out/bench_r02_${V}_$w.log | head -24
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02_${V}_$w.json'))
    print('$w', 'ms/step', round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['value'] * 1e3, 1), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'], 4),
          'int', round(d['roofline'].get('int_pipe', {}).get('frac', 0), 3), 'solve', round(d['schur_solve']['device_ms'], 1))
except Exception as e:
    print('$w', 'FAILED', e)
PY
done
