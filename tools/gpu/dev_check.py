"""Development check on the GPU box: bit-exact parity of the whole step against the
CPU oracle at 768 bits (the only precision a `make NLS=14` build carries), on shapes
that exercise ragged tiles, m = 2/3 blocks and multi-level factorisations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import sdpb_b200  # noqa: E402

KEYS = ["X_chol", "Y_chol", "A_X_inv", "A_Y", "L", "P", "Q"]
CASES = [
    (768, [(1, 5), (2, 4), (3, 2)], 4),
    (768, [(1, 24), (1, 25), (1, 31)], 20),
    (768, [(2, 20), (1, 40), (1, 33)], 50),
    (768, [(1, 40), (2, 9)], 70),
]
for prec, shapes, N in CASES:
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=3)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    got = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k, got[k], want[k])
    ctx.set_concurrency(0)
    again = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k + " (single stream)", again[k], want[k])
    ctx.close()
    print("parity ok", prec, shapes, N, flush=True)
