# dev: step_length on the device (row N3) -- parity tests at the precisions of the development build
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "search_direction_bit_exact and (768 or 664)" 2>&1 | tail -25
timeout 900 python -m pytest tests/test_golden_trajectory.py -x -q -m gpu 2>&1 | tail -15
