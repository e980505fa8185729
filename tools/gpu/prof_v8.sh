set -x
export SDPB_B200_CONCURRENCY=0
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trsm_gemm_level -s 28 -c 7 -o /tmp/trsm_levels python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu -i /tmp/trsm_levels.ncu-rep --page raw --csv > gpurun_out/prof_r01_v8_trsm_gemm_levels_raw.csv
ncu -i /tmp/trsm_levels.ncu-rep --page source --csv --kernel-id :::6 > gpurun_out/prof_r01_v8_trsm_gemm_It6_source.csv 2>/dev/null || ncu -i /tmp/trsm_levels.ncu-rep --page source --csv > gpurun_out/prof_r01_v8_trsm_gemm_source_all.csv
timeout 900 ncu --set full --clock-control none -k regex:solve_ -s 14 -c 7 -o /tmp/solve python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu -i /tmp/solve.ncu-rep --page raw --csv > gpurun_out/prof_r01_v8_solve_raw.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01_v8.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out /tmp/*.ncu-rep | tail -12
