"""Seeded synthetic block SDPs in the packed element format of the C-ABI
(SURVEY.md §8d): bilinear bases and free-variable matrices B with entries
U(-1,1) and full-length mantissas (no trailing zero limbs, which would flatter
every multiply), X and Y symmetric and strictly diagonally dominant.

Pure numpy on purpose: bench.py's product arm must not touch oracle/.
"""
import numpy as np

from .capi import BlockShape, elem_words, stored_limbs

# named workloads: (precision, [(m, n) ...], N).  C2-C4 are shape-matched
# stand-ins for the physics SDPs BASELINE.json names (no PMP inputs ship in the
# reference); C1's shapes are those of the shipped J=11 fixture.
WORKLOADS = {
    "c1": (768, [(1, n) for n in (24, 25, 26, 27, 28, 29, 30, 31, 31, 31, 30)], 20),
    "c2": (448, [(1, 30)] * 200, 60),
    "c3": (768, [(2, 40)] * 150 + [(1, 40)] * 450, 300),
    "c3-sample": (768, [(2, 40)] * 6 + [(1, 40)] * 18, 300),
    "c4": (960, [(2, 40)] * 250 + [(1, 40)] * 750, 1000),
    "c4-sample": (960, [(2, 40)] * 6 + [(1, 40)] * 18, 1000),
    "tiny": (768, [(1, 6), (2, 4), (1, 9)], 5),
    # C5, the synthetic sweep of BASELINE.json configs[4]: J in {256, 1024, 4096}, block dim P_j in
    # {64, 256} (m = 1, n = P_j), N in {512, 4096}, prec in {256, 768, 1536}.  Named corners; every
    # one needs sum P_j >= N (Q positive definite).  What does not fit 180 GB is listed in DESIGN.md.
    "c5-j256-p64-n512": (768, [(1, 64)] * 256, 512),
    "c5-j1024-p64-n512": (768, [(1, 64)] * 1024, 512),
    "c5-j4096-p64-n512": (768, [(1, 64)] * 4096, 512),
    "c5-j256-p256-n512": (768, [(1, 256)] * 256, 512),
    "c5-j1024-p256-n512": (768, [(1, 256)] * 1024, 512),
    "c5-j256-p64-n4096-256b": (256, [(1, 64)] * 256, 4096),
    "c5-j256-p64-n4096": (768, [(1, 64)] * 256, 4096),
    "c5-j256-p64-n4096-1536b": (1536, [(1, 64)] * 256, 4096),
    "c5-j256-p64-n512-1536b": (1536, [(1, 64)] * 256, 512),
    "c5-j256-p256-n512-256b": (256, [(1, 256)] * 256, 512),
    "c5-j256-p256-n512-1536b": (1536, [(1, 256)] * 256, 512),
    # bounded samples for the CPU restatement (same block shapes and N; rows >= N)
    "c5-j256-p64-n512-sample": (768, [(1, 64)] * 10, 512),
    "c5-j1024-p64-n512-sample": (768, [(1, 64)] * 10, 512),
    "c5-j4096-p64-n512-sample": (768, [(1, 64)] * 10, 512),
    "c5-j256-p256-n512-sample": (768, [(1, 256)] * 3, 512),
    "c5-j1024-p256-n512-sample": (768, [(1, 256)] * 3, 512),
    "c5-j256-p256-n512-256b-sample": (256, [(1, 256)] * 3, 512),
    "c5-j256-p256-n512-1536b-sample": (1536, [(1, 256)] * 3, 512),
}


def _header(exp, sign):
    return np.uint64((exp & 0xFFFFFFFF) | ((sign & 0xFFFFFFFF) << 32))


def random_matrix(rng, prec, h, w):
    """h x w, entries uniform in (-1,1): exponent 0, NL random limbs."""
    nl, ew = stored_limbs(prec), elem_words(prec)
    out = np.zeros((w, h, ew), dtype=np.uint64)
    if h * w == 0:
        return out
    out[:, :, 1:1 + nl] = rng.integers(0, 1 << 64, size=(w, h, nl), dtype=np.uint64)
    out[:, :, nl] |= np.uint64(1)  # top limb non-zero
    sign = rng.integers(0, 2, size=(w, h), dtype=np.uint64)
    out[:, :, 0] = np.where(sign == 1, _header(0, 1), _header(0, -1))
    return out


def random_spd(rng, prec, s):
    """Symmetric s x s: off-diagonal U(-1,1), diagonal in [s+1, s+2)."""
    nl, ew = stored_limbs(prec), elem_words(prec)
    a = random_matrix(rng, prec, s, s)
    if s == 0:
        return a
    iu = np.triu_indices(s, 1)
    a[iu[0], iu[1]] = a[iu[1], iu[0]]  # mirror (col,row) <- (row,col)
    d = np.arange(s)
    a[d, d, 0] = _header(1, 1)
    a[d, d, nl] = np.uint64(s + 1)
    assert ew >= nl + 1
    return a


class SyntheticSDP:
    """One rank's share of a synthetic block SDP, ready for any context that
    exposes set_block / schur_step (the CUDA library, or the oracle in tests)."""

    def __init__(self, prec, shapes, N, seed=1):
        self.prec, self.N = prec, N
        self.shapes = [s if isinstance(s, BlockShape) else BlockShape(*s) for s in shapes]
        rng = np.random.Generator(np.random.PCG64(0x5D9B0000 + seed))
        self.B, self.bases, self.X, self.Y = [], [], [], []
        for s in self.shapes:
            self.B.append(random_matrix(rng, prec, s.schur_size, N))
            self.bases.append((random_matrix(rng, prec, s.basis_height(0), s.n),
                               random_matrix(rng, prec, s.basis_height(1), s.n)))
            for p in (0, 1):
                self.X.append(random_spd(rng, prec, s.psd_size(p)))
                self.Y.append(random_spd(rng, prec, s.psd_size(p)))

    def upload(self, ctx):
        for j in range(len(self.shapes)):
            ctx.set_block(j, self.B[j], self.bases[j][0], self.bases[j][1])

    def input_bytes(self):
        return sum(a.nbytes for a in self.X) + sum(a.nbytes for a in self.Y)


def solve_rhs(prec, shapes, N, seed=7):
    """Right-hand sides (r_x per block, r_y) for solve_schur_complement_equation, U(-1,1)."""
    rng = np.random.Generator(np.random.PCG64(0x5D9B1000 + seed))
    shapes = [s if isinstance(s, BlockShape) else BlockShape(*s) for s in shapes]
    return [random_matrix(rng, prec, s.schur_size, 1) for s in shapes], random_matrix(rng, prec, N, 1)
