"""dev: exact syrk, integer tensor path against the IMAD.WIDE path on the same inputs (Q must agree bit for bit)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle_lib as ol
import sdpb_b200
from sdpb_b200.capi import SdpbB200Error

MODES = tuple(os.environ.get("MODES", "imad,imma").split(","))
CASES = [
    (768, [(1, 5), (2, 4), (3, 2)], 4),
    (768, [(1, 24), (1, 25), (1, 31)], 20),
    (768, [(1, 6), (2, 4), (1, 9)], 5),
    (768, [(1, 7), (2, 5), (1, 9), (2, 3)], 9),
    (768, [(2, 40), (1, 40), (1, 33)], 37),
    (768, [(2, 40)] * 3, 70),
]
for prec, shapes, N in CASES:
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=3)
    out = {}
    for mode in MODES:
        os.environ["SDPB_B200_SYRK"] = mode
        ctx = sdpb_b200.SchurContext(prec, shapes, N)
        sdp.upload(ctx)
        try:
            out[mode] = sdp.run_step(ctx)["Q"]
        except SdpbB200Error as e:
            out[mode] = str(e)
        ctx.close()
    a, b = out["imad"], out["imma"]
    K = sum(n * m * (m + 1) // 2 for m, n in shapes)
    if isinstance(b, str) or isinstance(a, str):
        print(prec, shapes, "N", N, "K", K, "->", a if isinstance(a, str) else "ok", "|", b if isinstance(b, str) else "ok")
        continue
    bad = np.argwhere((a != b).any(axis=-1))
    print(prec, shapes, "N", N, "K", K, "differing Q entries:", len(bad), bad[:6].tolist())
