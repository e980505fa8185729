# quick development session on a 768-bit-only build (make NLS=14): subset of parity tests + bench kernels table
set -x
export SDPB_B200_JIT=0
( time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharded_local.py -m gpu -x -q -k "${K:-c3_sample or wide_Q or diagonals or reduced or two_ranks or c3_sample_shapes or panel_distributed or failure_on_one or resident_step}" ) 2>&1 | tail -15
timeout 600 python bench.py --kernels --steps 5 --warmup 3 --no-cpu --no-all-outputs > gpurun_out/dev_bench.json 2> gpurun_out/dev_bench.log
grep -v "^\s*$" gpurun_out/dev_bench.log | head -60
python - <<'PY'
import json
d = json.load(open('gpurun_out/dev_bench.json'))
print({k: d[k] for k in ('ms_per_step', 'serial_ms_per_step', 'gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['int_pipe']['frac'])
PY
if [ -n "$TRSM_AB" ]; then
  SDPB_B200_TRSM=levels timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu --no-all-outputs 2>&1 >/dev/null | grep -E "trsm|stages"
fi
