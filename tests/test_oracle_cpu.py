"""Sanity of the CPU oracle itself (no GPU): factorisations reproduce their
inputs, the Schur step runs on small synthetic SDPs, error paths fire."""
import numpy as np
import pytest

import oracle_lib as ol
from sdpb_b200.capi import elem_words


def _to_float(prec, a):
    w, h, _ = a.shape
    out = np.zeros((h, w))
    for j in range(w):
        for i in range(h):
            out[i, j] = ol.to_double(prec, a[j, i])
    return out


@pytest.mark.parametrize("prec", [128, 768])
def test_cholesky_and_trsm_reproduce_inputs(prec):
    s, ncols = 7, 3
    A = ol.random_spd(prec, s, 5)
    L = np.zeros_like(A)
    lib = ol.load_oracle()
    assert lib.oracle_potrf(prec, s, 0, ol._ptr(A), ol._ptr(L)) == -1
    Lf, Af = _to_float(prec, L), _to_float(prec, A)
    assert np.allclose(Lf @ Lf.T, Af, rtol=1e-13, atol=1e-13)
    assert np.allclose(np.triu(Lf, 1), 0)
    U = np.zeros_like(A)
    assert lib.oracle_potrf(prec, s, 1, ol._ptr(A), ol._ptr(U)) == -1
    Uf = _to_float(prec, U)
    assert np.allclose(Uf.T @ Uf, Af, rtol=1e-13, atol=1e-13)
    B = ol.random_matrix(prec, s, ncols, 9)
    X = np.zeros_like(B)
    lib.oracle_trsm(prec, s, ncols, ol._ptr(L), ol._ptr(B), ol._ptr(X))
    assert np.allclose(Lf @ _to_float(prec, X), _to_float(prec, B), rtol=1e-12, atol=1e-12)


def test_non_positive_pivot_is_reported():
    prec, s = 256, 4
    A = ol.scale_matrix(prec, ol.random_spd(prec, s, 3), -1.0)
    L = np.zeros_like(A)
    assert ol.load_oracle().oracle_potrf(prec, s, 0, ol._ptr(A), ol._ptr(L)) == 0


@pytest.mark.parametrize("prec,shapes,N", [
    (128, [(1, 4), (2, 3), (1, 1)], 3),
    (768, [(1, 5), (2, 4)], 4),
])
def test_schur_step_is_consistent_in_double(prec, shapes, N):
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=7)
    ctx = ol.OracleContext(prec, shapes, N)
    sdp.upload(ctx)
    out = sdp.run_step(ctx)
    # Q = U^T U must equal sum_j B_j^T S_j^{-1} B_j ; check through P: Q = sum P_j^T P_j
    Uf = _to_float(prec, out["Q"])
    Q = Uf.T @ Uf
    acc = np.zeros((N, N))
    for j in range(len(shapes)):
        Pf = _to_float(prec, out["P"][j])
        acc += Pf.T @ Pf
    assert np.allclose(Q, acc, rtol=1e-10, atol=1e-10)
    # P_j = L_j^{-1} B_j
    for j in range(len(shapes)):
        Lf = _to_float(prec, out["L"][j])
        assert np.allclose(Lf @ _to_float(prec, out["P"][j]), _to_float(prec, sdp.B[j]), rtol=1e-9, atol=1e-9)
    # pairings are symmetric bit for bit
    for a in out["A_X_inv"] + out["A_Y"]:
        assert np.array_equal(a, a.transpose(1, 0, 2))


def test_non_pd_X_names_block_and_parity():
    prec, shapes, N = 128, [(1, 4), (1, 3)], 2
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=2)
    sdp.X[3] = ol.scale_matrix(prec, sdp.X[3], -1.0)
    ctx = ol.OracleContext(prec, shapes, N)
    sdp.upload(ctx)
    with pytest.raises(ol.OracleError) as ei:
        sdp.run_step(ctx)
    assert "block index = 1, parity = 1" in str(ei.value)


@pytest.mark.parametrize("prec,shapes,N", [
    (128, [(1, 4), (2, 3), (1, 1)], 3),
    (768, [(1, 5), (2, 4), (1, 7)], 4),
])
def test_schur_solve_satisfies_the_block_system(prec, shapes, N):
    """solve_schur_complement_equation.cxx:16-79 solves {{S, -B}, {B^T, 0}} {dx, dy} = {r_x, r_y}
    (compute_schur_RHS.cxx:3-7): check S_j dx_j - B_j dy = r_x_j and sum_j B_j^T dx_j = r_y in double."""
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=11)
    ctx = ol.OracleContext(prec, shapes, N)
    sdp.upload(ctx)
    out = sdp.run_step(ctx)
    rx, ry = sdp.solve_rhs()
    dx, dy = [a.copy() for a in rx], ry.copy()
    ctx.solve_schur_complement_equation(dx, dy)
    dyf = _to_float(prec, dy)
    acc = np.zeros((N, 1))
    for j in range(len(shapes)):
        Lf, Bf = _to_float(prec, out["L"][j]), _to_float(prec, sdp.B[j])
        dxf = _to_float(prec, dx[j])
        assert np.allclose(Lf @ (Lf.T @ dxf) - Bf @ dyf, _to_float(prec, rx[j]), rtol=1e-8, atol=1e-8)
        acc += Bf.T @ dxf
    assert np.allclose(acc, _to_float(prec, ry), rtol=1e-8, atol=1e-8)
    # the staged (sharded) model with one rank is the same computation
    dx2, dy2 = [a.copy() for a in rx], ry.copy()
    part = ctx.shard_solve_stage1(dx2)
    ctx.shard_solve_stage2(part, dx2, dy2)
    ol.assert_same("dx", dx2, dx)
    ol.assert_same("dy", dy2, dy)


def test_schur_solve_needs_the_factors():
    prec, shapes, N = 128, [(1, 4)], 2
    ctx = ol.OracleContext(prec, shapes, N)
    dx, dy = ctx.alloc_solve_vectors()
    with pytest.raises(ol.OracleError) as ei:
        ctx.solve_schur_complement_equation(dx, dy)
    assert ei.value.code == 5


def test_scale_multiply_add_in_double():
    """scale_multiply_add.cxx:4-16: C = alpha A B + beta C per block, the three (alpha, beta) the
    reference uses (step.cxx:137, compute_search_direction.cxx:28,60)."""
    prec, shapes, N = 256, [(1, 5), (2, 3), (1, 1)], 2
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=13)
    ctx = ol.OracleContext(prec, shapes, N)
    for alpha, beta in ((-1, 0), (1, 0), (-1, 1)):
        C = [a.copy() for a in sdp.Y]
        ctx.scale_multiply_add(alpha, sdp.X, sdp.Y, beta, C)
        for a, b, c0, c in zip(sdp.X, sdp.Y, sdp.Y, C):
            if a.size == 0:
                continue
            want = alpha * (_to_float(prec, a) @ _to_float(prec, b)) + beta * _to_float(prec, c0)
            assert np.allclose(_to_float(prec, c), want, rtol=1e-12, atol=1e-12)
    with pytest.raises(ol.OracleError):
        ctx.scale_multiply_add(2, sdp.X, sdp.Y, 0, [a.copy() for a in sdp.Y])


@pytest.mark.parametrize("prec,K,N", [(256, 37, 5), (768, 90, 11)])
def test_exact_syrk_direct_equals_the_references_crt_blas_route(prec, K, N):
    """Row a9: the oracle's direct mpz sum and the reference's formulation (residues modulo the
    primes of Fmpz_Comb.cxx:22-68, fp64 dsyrk per prime, CRT) give the same Q' to the last limb --
    the check the reference's calculate_matrix_square.test.cxx:44-86 makes against fmpz_mat_mul_blas."""
    Pn = ol.integer_valued_matrix(prec, K, N, seed=17)
    primes = ol.syrk_crt_primes(prec, K)
    assert all(int(p) < 1664544 for p in primes) and len(set(primes.tolist())) == len(primes)
    prod = 1
    for p in primes:
        prod *= int(p)
    assert prod.bit_length() > 2 * prec + K.bit_length() + 1
    want = ol.syrk_direct(prec, Pn)
    got = ol.syrk_crt_blas(prec, Pn)
    iu = np.triu_indices(N)
    assert np.array_equal(got[iu[1], iu[0]], want[iu[1], iu[0]])  # (col j, row i), i <= j
    assert want[iu[1], iu[0]].any()


def _pack_doubles(prec, M):
    s = M.shape[0]
    out = np.zeros((s, s, elem_words(prec)), dtype=np.uint64)
    for j in range(s):
        for i in range(s):
            out[j, i] = ol.from_decimal(prec, "%.40e" % M[i, j])
    return out


def _oracle_min_eigenvalue(prec, M, L=None):
    import ctypes
    s = M.shape[0]
    out = np.zeros(elem_words(prec), dtype=np.uint64)
    it = ctypes.c_int(0)
    A = _pack_doubles(prec, M)
    Lp = None if L is None else ol._ptr(_pack_doubles(prec, L))
    ol.load_oracle().oracle_min_eigenvalue(prec, s, Lp, ol._ptr(A), ol._ptr(out), ctypes.byref(it))
    return ol.to_double(prec, out), it.value


def test_min_eigenvalue_restatement_against_lapack():
    """Row N3's canonical algorithm (csrc/host/step_length.hpp: congruence, Householder
    tridiagonalisation, Laguerre's iteration from the Gershgorin bound) against LAPACK in double:
    generic matrices, the degenerate spectra that slow Laguerre down to linear convergence, exact
    zeros, and the lower triangle being the one that is read (min_eigenvalue.cxx:28, LOWER)."""
    prec = 768
    rng = np.random.default_rng(5)

    def check(M, L=None, max_iters=None):
        got, iters = _oracle_min_eigenvalue(prec, M, L)
        S = np.tril(M) + np.tril(M, -1).T
        if L is not None:
            Li = np.linalg.inv(L)
            S = Li @ M @ Li.T
            S = np.tril(S) + np.tril(S, -1).T
        ev = np.linalg.eigvalsh(S)
        scale = max(np.abs(ev).max(), 1e-300)
        assert abs(got - ev[0]) <= 1e-12 * scale, (got, ev[0], iters)
        if max_iters is not None:
            assert iters <= max_iters, iters
        return iters

    for n in (1, 2, 3, 5, 20, 40):
        A = rng.standard_normal((n, n))
        check(A + A.T, max_iters=12)
    check(np.diag(rng.standard_normal(10)), max_iters=2)
    check(np.zeros((6, 6)), max_iters=0)
    check(-3.0 * np.eye(7), max_iters=2)
    Q, _ = np.linalg.qr(rng.standard_normal((12, 12)))
    check(Q @ np.diag([1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11.]) @ Q.T, max_iters=80)      # double root (to 1e-16)
    check(Q @ np.diag([-1, -1, -1, 3, 4, 5, 6, 7, 8, 9, 10, 11.]) @ Q.T, max_iters=100)  # triple root
    check(Q @ np.diag([1.0] * 11 + [2.0]) @ Q.T, max_iters=100)
    W = np.diag(np.abs(np.arange(-10, 11)).astype(float)) + np.diag(np.ones(20), 1) + np.diag(np.ones(20), -1)
    check(-W, max_iters=60)  # Wilkinson's matrix: the two largest eigenvalues agree to 1e-15
    check(rng.standard_normal((9, 9)))  # not symmetric: only the lower triangle counts
    n = 30
    A = rng.standard_normal((n, n))
    B = rng.standard_normal((n, n))
    check(A + A.T, L=np.linalg.cholesky(B @ B.T + n * np.eye(n)), max_iters=12)
