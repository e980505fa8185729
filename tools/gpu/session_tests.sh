# full GPU test suite, verbose, whole log kept
set -x
( time timeout 1500 python -m pytest tests -m gpu -v -x ) > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "PASSED|FAILED|ERROR|passed|failed|Abort|abort|terminate" gpurun_out/pytest_gpu_full.log | tail -90 | cut -c1-160
