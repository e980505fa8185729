set -x
timeout 300 python tools/gpu/dev_coop.py 2>&1 | tail -8
timeout 900 python tools/gpu/dev_check.py 2>&1 | tail -8
timeout 600 python bench.py --kernels --steps 3 --warmup 3 --no-cpu 2>&1 | tail -25
SDPB_B200_CONCURRENCY=0 timeout 900 ncu --set full --clock-control none -k regex:"trsm_(gemm|diag)_level" -s 65 -c 15 --csv --page raw --log-file gpurun_out/traffic_trsm_r01_v5.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
