"""Test-side binding of the CPU oracle (oracle/liboracle.so) — same method names
as sdpb_b200.SchurContext so parity tests drive both with identical inputs.

The oracle is test infrastructure: nothing under sdpb_b200/ imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

from sdpb_b200.capi import StepContextBase, ptr_array, _ptr, u64p, u64pp, elem_words

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
_lib = None


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def load_oracle():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    lib = ctypes.CDLL(ORACLE_SO)
    lib.oracle_create.restype = ctypes.c_int
    lib.oracle_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.oracle_destroy.argtypes = [ctypes.c_void_p]
    lib.oracle_destroy.restype = None
    lib.oracle_last_error.restype = ctypes.c_char_p
    lib.oracle_qprime_words.restype = ctypes.c_int
    lib.oracle_qprime_words.argtypes = [ctypes.c_int]
    lib.oracle_shard_stage1.restype = ctypes.c_int
    lib.oracle_shard_stage1.argtypes = [ctypes.c_void_p, u64pp, u64pp, u64p]
    lib.oracle_shard_stage2.restype = ctypes.c_int
    lib.oracle_shard_stage2.argtypes = [ctypes.c_void_p, u64p, ctypes.c_int, u64p]
    lib.oracle_shard_stage3.restype = ctypes.c_int
    lib.oracle_shard_stage3.argtypes = [ctypes.c_void_p, ctypes.c_int, u64pp, u64pp, u64pp, u64p]
    lib.oracle_last_error.argtypes = [ctypes.c_void_p]
    lib.oracle_solve_schur_complement_equation.restype = ctypes.c_int
    lib.oracle_solve_schur_complement_equation.argtypes = [ctypes.c_void_p, u64pp, u64p]
    lib.oracle_syrk_crt_primes.restype = ctypes.c_int
    lib.oracle_syrk_crt_primes.argtypes = [ctypes.c_int, ctypes.c_long, u64p, ctypes.c_int]
    lib.oracle_syrk_crt_begin.restype = ctypes.c_void_p
    lib.oracle_syrk_crt_begin.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_int, u64p]
    lib.oracle_syrk_crt_residues.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double)]
    lib.oracle_syrk_crt_end.restype = None
    lib.oracle_syrk_crt_end.argtypes = [ctypes.c_void_p]
    lib.oracle_syrk_crt_reconstruct.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, u64p,
                                                ctypes.POINTER(ctypes.c_int64), u64p]
    lib.oracle_syrk_direct.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_int, u64p, u64p]
    lib.oracle_scale_multiply_add.restype = ctypes.c_int
    lib.oracle_scale_multiply_add.argtypes = [ctypes.c_void_p, ctypes.c_int, u64pp, u64pp, ctypes.c_int, u64pp]
    lib.oracle_shard_solve_stage1.restype = ctypes.c_int
    lib.oracle_shard_solve_stage1.argtypes = [ctypes.c_void_p, u64pp, u64p]
    lib.oracle_shard_solve_stage2.restype = ctypes.c_int
    lib.oracle_shard_solve_stage2.argtypes = [ctypes.c_void_p, u64p, ctypes.c_int, u64pp, u64p]
    lib.oracle_set_block.argtypes = [ctypes.c_void_p, ctypes.c_int, u64p, u64p, u64p]
    lib.oracle_cholesky_decomposition.argtypes = [ctypes.c_void_p, ctypes.c_int, u64pp, u64pp]
    lib.oracle_compute_bilinear_pairings.argtypes = [ctypes.c_void_p, u64pp, u64pp, u64pp]
    lib.oracle_initialize_schur_complement_solver.argtypes = [
        ctypes.c_void_p, u64pp, u64pp, u64p, ctypes.POINTER(ctypes.c_int32)]
    lib.oracle_schur_step.argtypes = [ctypes.c_void_p] + [u64pp] * 8 + [u64p, ctypes.POINTER(ctypes.c_int32)]
    lib.oracle_potrf.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, u64p, u64p]
    lib.oracle_trsm.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, u64p, u64p, u64p]
    lib.oracle_scalar_op.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long, u64p, u64p, u64p]
    lib.oracle_from_decimal.argtypes = [ctypes.c_int, ctypes.c_char_p, u64p]
    lib.oracle_to_double.restype = ctypes.c_double
    lib.oracle_to_double.argtypes = [ctypes.c_int, u64p]
    lib.oracle_random_matrix.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, u64p]
    lib.oracle_random_spd.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, u64p]
    lib.oracle_scale_matrix.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_double, u64p, u64p]
    lib.oracle_stage_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_int]
    lib.oracle_num_threads.restype = ctypes.c_int
    for name, args in (("direction_begin", [u64p]), ("direction_R_errors", [u64p, u64p]),
                       ("direction_set_residues", [u64pp, u64pp, u64p]),
                       ("compute_search_direction", [u64p, ctypes.c_int]), ("direction_frobenius", [u64p]),
                       ("direction_get", [u64pp, u64pp, u64p, u64pp]), ("step_length", [ctypes.c_int, u64p]),
                       ("direction_put", [u64pp, u64pp])):
        f = getattr(lib, "oracle_" + name)
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p] + args
    lib.oracle_min_eigenvalue.restype = ctypes.c_int
    lib.oracle_min_eigenvalue.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, ctypes.POINTER(ctypes.c_int)]
    lib.oracle_set_num_threads.restype = None
    lib.oracle_set_num_threads.argtypes = [ctypes.c_int]
    _lib = lib
    return lib


class OracleError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"oracle error {code}: {message}")
        self.code = code
        self.message = message


class OracleContext(StepContextBase):
    PREFIX = "oracle_"  # StepContextBase's direction_* methods call oracle_direction_* here

    def __init__(self, prec_bits, shapes, N):
        super().__init__(prec_bits, shapes, N)
        self.lib = load_oracle()
        self.handle = ctypes.c_void_p()
        dims = (ctypes.c_int * max(1, self.J))(*[s.m for s in self.shapes])
        npts = (ctypes.c_int * max(1, self.J))(*[s.n for s in self.shapes])
        self.lib.oracle_create(ctypes.byref(self.handle), prec_bits, self.J, dims, npts, N)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.oracle_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OracleError(rc, self.lib.oracle_last_error(self.handle).decode())

    def set_block(self, j, B, bases_even, bases_odd):
        self._check(self.lib.oracle_set_block(self.handle, j, _ptr(B), _ptr(bases_even), _ptr(bases_odd)))

    def cholesky_decomposition(self, which, A, L=None):
        self._check(self.lib.oracle_cholesky_decomposition(
            self.handle, which, ptr_array(A), ptr_array(L) if L is not None else None))

    def compute_bilinear_pairings(self, Y, A_X_inv=None, A_Y=None):
        self._check(self.lib.oracle_compute_bilinear_pairings(
            self.handle, ptr_array(Y),
            ptr_array(A_X_inv) if A_X_inv is not None else None,
            ptr_array(A_Y) if A_Y is not None else None))

    def initialize_schur_complement_solver(self, L=None, P=None, Q=None, block_timings_ms=None):
        self._check(self.lib.oracle_initialize_schur_complement_solver(
            self.handle,
            ptr_array(L) if L is not None else None,
            ptr_array(P) if P is not None else None,
            _ptr(Q) if Q is not None else None, None))

    def schur_step(self, X, Y, X_chol=None, Y_chol=None, A_X_inv=None, A_Y=None, L=None, P=None, Q=None,
                   block_timings_ms=None):
        opt = lambda v: ptr_array(v) if v is not None else None  # noqa: E731
        self._check(self.lib.oracle_schur_step(
            self.handle, ptr_array(X), ptr_array(Y), opt(X_chol), opt(Y_chol), opt(A_X_inv), opt(A_Y),
            opt(L), opt(P), _ptr(Q) if Q is not None else None, None))

    def solve_schur_complement_equation(self, dx, dy):
        self._check(self.lib.oracle_solve_schur_complement_equation(self.handle, ptr_array(dx), _ptr(dy)))

    def scale_multiply_add(self, alpha, A, B, beta, C):
        self._check(self.lib.oracle_scale_multiply_add(self.handle, int(alpha), ptr_array(A), ptr_array(B),
                                                       int(beta), ptr_array(C)))

    # sharded Schur solve: stage 1 local (dx in place, partial rows out), the caller gathers the
    # rows in global block order, stage 2 finishes dy (replicated) and the local dx
    def shard_solve_stage1(self, dx):
        part = np.zeros((self.J, self.N, self.ew), dtype=np.uint64)
        self._check(self.lib.oracle_shard_solve_stage1(self.handle, ptr_array(dx), _ptr(part)))
        return part

    def shard_solve_stage2(self, part_global, dx, dy):
        part_global = np.ascontiguousarray(part_global, dtype=np.uint64)
        self._check(self.lib.oracle_shard_solve_stage2(self.handle, _ptr(part_global), part_global.shape[0],
                                                       ptr_array(dx), _ptr(dy)))

    # sharded model (oracle_shard_stage1..3): the caller does the two exchanges
    def shard_stage1(self, X, Y):
        part = np.zeros((self.J, self.N, self.ew), dtype=np.uint64)
        self._check(self.lib.oracle_shard_stage1(self.handle, ptr_array(X), ptr_array(Y), _ptr(part)))
        return part

    def shard_stage2(self, part_global):
        part_global = np.ascontiguousarray(part_global, dtype=np.uint64)
        W = self.lib.oracle_qprime_words(self.prec)
        q = np.zeros((self.N, self.N, W), dtype=np.uint64)
        self._check(self.lib.oracle_shard_stage2(self.handle, _ptr(part_global), part_global.shape[0], _ptr(q)))
        return q

    def shard_stage3(self, qparts):
        L, P, Q = self.alloc_schur_outputs()
        qparts = [np.ascontiguousarray(q, dtype=np.uint64) for q in qparts]
        self._check(self.lib.oracle_shard_stage3(self.handle, len(qparts), ptr_array(qparts), ptr_array(L),
                                                 ptr_array(P), _ptr(Q)))
        return L, P, Q

    def stage_ms(self):
        ms = (ctypes.c_double * 9)()
        self.lib.oracle_stage_ms(self.handle, ms, 9)
        return list(ms)


# ---------------------------------------------------------------- generators
def random_matrix(prec, h, w, seed):
    out = np.zeros((w, h, elem_words(prec)), dtype=np.uint64)
    if h * w:
        load_oracle().oracle_random_matrix(prec, h, w, seed, _ptr(out))
    return out


def random_spd(prec, s, seed):
    out = np.zeros((s, s, elem_words(prec)), dtype=np.uint64)
    if s:
        load_oracle().oracle_random_spd(prec, s, seed, _ptr(out))
    return out


def scale_matrix(prec, a, scale):
    out = np.zeros_like(a)
    if a.size:
        load_oracle().oracle_scale_matrix(prec, a.size // elem_words(prec), float(scale), _ptr(a), _ptr(out))
    return out


def from_decimal(prec, text):
    out = np.zeros(elem_words(prec), dtype=np.uint64)
    rc = load_oracle().oracle_from_decimal(prec, text.encode(), _ptr(out))
    if rc:
        raise ValueError(f"cannot parse {text!r}")
    return out


def to_double(prec, elem):
    e = np.ascontiguousarray(elem, dtype=np.uint64)
    return load_oracle().oracle_to_double(prec, _ptr(e))


def scalar_op(prec, op, a, b, k=0):
    r = np.zeros_like(a)
    load_oracle().oracle_scalar_op(prec, op, k, a.size // elem_words(prec), _ptr(a), _ptr(b), _ptr(r))
    return r


def syrk_crt_primes(prec, k):
    buf = np.zeros(4096, dtype=np.uint64)
    n = load_oracle().oracle_syrk_crt_primes(prec, k, _ptr(buf), len(buf))
    assert n > 0
    return buf[:n].copy()


def syrk_direct(prec, Pn):
    """Q' = trunc(P')^T trunc(P') by the direct mpz sum; Pn: (N, K, ew) packed, integer-valued."""
    N, K, _ = Pn.shape
    Q = np.zeros((N, N, elem_words(prec)), dtype=np.uint64)
    load_oracle().oracle_syrk_direct(prec, K, N, _ptr(Pn), _ptr(Q))
    return Q


def syrk_crt_blas(prec, Pn, timings=None):
    """The same Q' the way the reference computes it (bigint_syrk/Readme.md:27-55): residues modulo
    the primes of Fmpz_Comb.cxx:22-68 as symmetric doubles, one fp64 dsyrk per prime (scipy's
    OpenBLAS), CRT.  timings: optional dict receiving seconds per phase."""
    import time
    from scipy.linalg.blas import dsyrk
    lib = load_oracle()
    N, K, _ = Pn.shape
    primes = syrk_crt_primes(prec, K)
    t0 = time.perf_counter()
    st = lib.oracle_syrk_crt_begin(prec, K, N, _ptr(Pn))
    t_conv = time.perf_counter() - t0
    res = np.zeros((len(primes), N, N), dtype=np.int64)
    A = np.zeros((K, N), dtype=np.float64, order="F")
    t_res = t_blas = 0.0
    for k, p in enumerate(primes):
        t0 = time.perf_counter()
        lib.oracle_syrk_crt_residues(st, int(p), A.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        t1 = time.perf_counter()
        C = dsyrk(1.0, A, trans=1, lower=0)             # upper triangle of A^T A, exact in fp64
        res[k] = np.mod(C.T.astype(np.int64), int(p))   # res[k][j][i] = C(i, j): column-major (i + j N)
        t2 = time.perf_counter()
        t_res += t1 - t0
        t_blas += t2 - t1
    lib.oracle_syrk_crt_end(st)
    t0 = time.perf_counter()
    Q = np.zeros((N, N, elem_words(prec)), dtype=np.uint64)
    lib.oracle_syrk_crt_reconstruct(prec, N, len(primes), _ptr(primes),
                                    res.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _ptr(Q))
    t_crt = time.perf_counter() - t0
    if timings is not None:
        timings.update(convert=t_conv, residues=t_res, dsyrk_and_mod=t_blas, crt=t_crt, primes=len(primes))
    return Q


def integer_valued_matrix(prec, K, N, seed):
    """K x N with entries uniform in (-2^prec, 2^prec): what normalize_and_shift hands to the syrk."""
    a = random_matrix(prec, K, N, seed)
    return scalar_op(prec, 5, a.reshape(-1, elem_words(prec)), a.reshape(-1, elem_words(prec)), prec).reshape(a.shape)


class SyntheticSDP:
    """Seeded synthetic block SDP of a given shape: bilinear
    bases and B with entries U(-1,1), full-length mantissas; X, Y symmetric
    positive definite.  The same object feeds the oracle and the CUDA library."""

    def __init__(self, prec, shapes, N, seed=1, block_ids=None):
        """block_ids: global indices of `shapes` when this object holds one rank's share of a
        sharded SDP (the data of a block depends only on its global index)."""
        from sdpb_b200.capi import BlockShape
        self.prec = prec
        self.shapes = [s if isinstance(s, BlockShape) else BlockShape(*s) for s in shapes]
        self.N = N
        self.B, self.bases = [], []
        self.X, self.Y = [], []
        ids = list(block_ids) if block_ids is not None else list(range(len(self.shapes)))
        self.block_ids = ids
        for j, s in zip(ids, self.shapes):
            base = seed * 1000003 + j * 101
            self.B.append(random_matrix(prec, s.schur_size, N, base + 1))
            self.bases.append((random_matrix(prec, s.basis_height(0), s.n, base + 2),
                               random_matrix(prec, s.basis_height(1), s.n, base + 3)))
            for p in (0, 1):
                self.X.append(random_spd(prec, s.psd_size(p), base + 10 + p))
                self.Y.append(random_spd(prec, s.psd_size(p), base + 20 + p))

    def solve_rhs(self, seed=7):
        """Seeded right-hand sides (r_x per block, r_y) for solve_schur_complement_equation; the
        data of a block depends only on its global index."""
        ids = self.block_ids
        dx = [random_matrix(self.prec, s.schur_size, 1, seed * 7919 + 31 * j + 5) for j, s in zip(ids, self.shapes)]
        dy = random_matrix(self.prec, self.N, 1, seed * 7919 + 3)
        return dx, dy

    def upload(self, ctx):
        for j in range(len(self.shapes)):
            ctx.set_block(j, self.B[j], self.bases[j][0], self.bases[j][1])

    def run_step(self, ctx):
        """Full hot path through `ctx`; returns every output as numpy arrays."""
        out = {
            "X_chol": ctx.alloc_psd_blocks(), "Y_chol": ctx.alloc_psd_blocks(),
            "A_X_inv": ctx.alloc_pairing_blocks(), "A_Y": ctx.alloc_pairing_blocks(),
        }
        out["L"], out["P"], out["Q"] = ctx.alloc_schur_outputs()
        ctx.schur_step(self.X, self.Y, out["X_chol"], out["Y_chol"], out["A_X_inv"], out["A_Y"],
                       out["L"], out["P"], out["Q"])
        return out


def assert_same(name, got, want):
    """Bit-exact comparison of packed arrays (lists of arrays or arrays)."""
    if isinstance(got, list):
        assert len(got) == len(want), name
        for i, (g, w) in enumerate(zip(got, want)):
            assert_same(f"{name}[{i}]", g, w)
        return
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    if not np.array_equal(got, want):
        bad = np.argwhere((got != want).any(axis=-1))
        first = tuple(bad[0])
        raise AssertionError(
            f"{name}: {len(bad)} of {got.shape[0] * got.shape[1]} elements differ; first at "
            f"(col,row)={first}: got {[hex(int(x)) for x in got[first]]} want "
            f"{[hex(int(x)) for x in want[first]]}")
