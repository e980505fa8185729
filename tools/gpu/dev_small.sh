set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "128 or 256 or 448 or c2" 2>&1 | tail -12
