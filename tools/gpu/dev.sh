set -x
timeout 300 python tools/gpu/dev_coop.py 2>&1 | tail -8
timeout 900 python tools/gpu/dev_check.py 2>&1 | tail -8
timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu 2>&1 | tail -25
