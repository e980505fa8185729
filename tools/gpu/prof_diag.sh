set -x
export SDPB_B200_GROUPS=1 SDPB_B200_CONCURRENCY=0
timeout 900 ncu --set full --import-source on --clock-control none -k regex:potrf_diag_rl -s 25 -c 1 -o gpurun_out/prof_r01_v5_diag_rl python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:potrf_diag_level -s 30 -c 1 -o gpurun_out/prof_r01_v5_diag_level python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
