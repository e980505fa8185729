# compute-sanitizer over the small end-to-end paths: memcheck (global/shared out-of-bounds, misaligned),
# racecheck (shared-memory hazards of the new kernels) and initcheck on smoke()
set -x
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_smoke.log 2>&1
  tail -6 gpurun_out/sanitizer_${tool}_smoke.log | cut -c1-200
done
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -k "step_length_degenerate or (block_groups and size) or (search_direction_bit_exact and 664)" > gpurun_out/sanitizer_memcheck_tests.log 2>&1
tail -8 gpurun_out/sanitizer_memcheck_tests.log | cut -c1-200
