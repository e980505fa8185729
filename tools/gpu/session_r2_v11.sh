# round-2 v11 (1 GPU, final tree of the round): all GPU tests, smoke, the default bench and the reference
# arm, c1 / c2, the ncu launch list of the bench command, DRAM traffic of the trsm stage
set -x
V=${1:-v11}
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02_$V.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
grep -v "^\s*$" gpurun_out/bench_r02_$V.log | head -48
python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_$V.json'))
print('$V', d['ms_per_step'], d['e2e'], d['stages_ms'], d['search_direction']['device_ms'], d['step_length']['device_ms'], d['e2e_newton_iteration']['value'], d['roofline']['frac'], d['roofline']['int_pipe']['frac'])
PY
for w in c1 c2; do
  timeout 600 python bench.py --workload $w --kernels --steps 10 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/bench_r02_${V}_$w.json 2> gpurun_out/bench_r02_${V}_$w.log
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_r02_${V}_$w.json'))
print('$w', 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'] * 1e3, 2), d['stages_ms'], 'newton', d['e2e_newton_iteration']['value'])
PY
done
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_${V}_ref.json 2> gpurun_out/bench_r02_${V}_ref.log ) 2>&1 | tail -3
cat gpurun_out/bench_r02_${V}_ref.json | cut -c1-400
SDPB_B200_CONCURRENCY=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_r02_$V.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"trsm_" -c 120 --csv --log-file gpurun_out/traffic_r02_$V.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-all-outputs > /dev/null 2>&1
ls -la gpurun_out | tail -5
