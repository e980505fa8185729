# final tree, 1 GPU: whole GPU test suite, smoke, default bench line, the 256-bit corner, block-group experiment
set -x
V=${1:-v6}
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_$V.json'))
print('$V', d['ms_per_step'], d['e2e'], d['cpu_baseline']['value'], d['stages_ms'])
PY
( time timeout 900 python bench.py --workload c5-j256-p64-n4096-256b --kernels --steps 2 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_${V}_c5-j256-p64-n4096-256b.json 2> gpurun_out/bench_r02_${V}_c5-j256-p64-n4096-256b.log ) 2>&1 | grep real
for g in 2 3; do
  SDPB_B200_GROUPS=$g timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_groups$g.json 2> /dev/null
  python - <<PY
import json
d = json.load(open('gpurun_out/dev_groups$g.json'))
print('groups $g', d['ms_per_step'], d['e2e']['value'])
PY
done
