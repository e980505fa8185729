set -x
SDPB_B200_CONCURRENCY=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r01_v10.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out/launches_r01_v10.csv
