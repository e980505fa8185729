set -x
timeout 600 python bench.py --kernels --steps 5 --warmup 3 --no-cpu 2>&1 | grep -v "^\s*$" | head -48
