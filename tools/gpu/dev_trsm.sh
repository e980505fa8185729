set -x
export SDPB_B200_JIT=0
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "c3_sample or resident_step" ) 2>&1 | tail -3
for m in levels levels1 hybrid; do
  SDPB_B200_TRSM=$m timeout 600 python bench.py --kernels --steps 4 --warmup 3 --no-cpu --no-all-outputs 2>&1 >gpurun_out/dev_trsm_$m.json | grep -E "trsm|stages"
  python -c "
import json; d=json.load(open('gpurun_out/dev_trsm_$m.json')); print('$m', d['ms_per_step'], d['serial_ms_per_step'])"
done
