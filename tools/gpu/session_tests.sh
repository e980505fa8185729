# full GPU test suite, whole log kept
set -x
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_full.log 2>&1
tail -12 gpurun_out/pytest_gpu_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
