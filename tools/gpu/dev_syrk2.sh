set -x
MODES=imma,imad timeout 300 python tools/gpu/dev_syrk_cmp.py
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "schur_step_bit_exact and (768 or 664)" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "768 or 664 or c3" 2>&1 | tail -12
