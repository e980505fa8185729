"""Dev check: the warp-cooperative pivot (sqrt + reciprocal) against libgmp / the single-thread routines."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
import sdpb_b200
from sdpb_b200.capi import elem_words

def operands(prec, count, seed):
    ew, nl = elem_words(prec), (prec + 63) // 64 + 2
    rng = np.random.default_rng(seed)
    a = ol.random_matrix(prec, count, 1, seed).reshape(count, ew).copy()
    for i in range(count):
        mode = i % 6
        if mode == 1:
            a[i, 1:nl + 1] = np.uint64(0xFFFFFFFFFFFFFFFF)
        elif mode == 2:
            a[i, 1:nl] = 0; a[i, nl] = 1
        elif mode == 3:
            a[i, 1:nl - 1] = 0
        elif mode == 4:
            a[i, nl] = np.uint64(int(rng.integers(1, 4)))
        if a[i, nl] == 0:
            a[i, nl] = 1
        e = int(rng.integers(-5, 6))
        a[i, 0] = np.uint64((e & 0xFFFFFFFF) | (1 << 32))
    return a

for prec in [int(x) for x in sys.argv[1:]] or [768]:
    ctx = sdpb_b200.SchurContext(prec, [(1, 2)], 1)
    a = operands(prec, 4096, 7)
    got = ctx.scalar_op(8, a, a)
    want = ol.scalar_op(prec, 4, a, a)
    bad = np.argwhere((got != want).any(axis=1))
    print(prec, "coop sqrt mismatches:", len(bad), bad[:5].ravel())
    got = ctx.scalar_op(9, a, a)
    cnt = got[:, 1] & np.uint64(0xFFFFFFFF)
    print(prec, "coop reciprocal: elements with mismatching words:", int((cnt != 0).sum()), "max", int(cnt.max()))
    ctx.close()

ctx = sdpb_b200.SchurContext(768, [(1, 2)], 1)
a = operands(768, 4096, 7)
for cnt in (1, 1, 148, 4096):
    ctx.scalar_op(8, a[:cnt].copy(), a[:cnt].copy())
    print("count", cnt, ctx.kernel_timings())
ctx.close()
