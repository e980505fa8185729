# round-2 GPU session: parity tests, the default bench line, the reference arm.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu/session_r2.sh v1'
set -x
V=${1:-v1}
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
nproc
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r02_$V.log
timeout 600 python bench.py --kernels --steps 5 --warmup 3 > gpurun_out/bench_r02_$V.json 2> gpurun_out/bench_r02_$V.log
tail -40 gpurun_out/bench_r02_$V.log
cat gpurun_out/bench_r02_$V.json
if [ "${2:-ref}" = "ref" ]; then
  ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_${V}_ref.json 2> gpurun_out/bench_r02_${V}_ref.log ) 2>&1 | tail -3
  tail -5 gpurun_out/bench_r02_${V}_ref.log
  cat gpurun_out/bench_r02_${V}_ref.json
fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
