set -x
./build/mac_bench_base quick
./build/mac_bench_one quick
timeout 600 python bench.py --kernels --steps 5 --warmup 3 --no-cpu 2>&1 | grep -v "^\s*$" | head -45
