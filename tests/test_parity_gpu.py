"""GPU parity: the CUDA library (through its C-ABI) against the CPU oracle on
identical seeded inputs.  Bit-exact: every limb, sign and exponent of every
output of the hot path."""
import numpy as np
import pytest

import oracle_lib as ol
import sdpb_b200
from sdpb_b200.capi import elem_words

pytestmark = pytest.mark.gpu

KEYS = ["X_chol", "Y_chol", "A_X_inv", "A_Y", "L", "P", "Q"]


def _adversarial_operands(prec, count, seed):
    """Random operands plus cancellation-heavy pairs for the scalar parity."""
    ew, nl = elem_words(prec), (prec + 63) // 64 + 2
    rng = np.random.default_rng(seed)
    a = ol.random_matrix(prec, count, 1, seed).reshape(count, ew).copy()
    b = ol.random_matrix(prec, count, 1, seed + 1).reshape(count, ew).copy()
    for i in range(count):
        mode = i % 8
        ea, eb = int(rng.integers(-3, 4)), int(rng.integers(-3, 4))
        sa, sb = (1 if rng.integers(2) else -1), (1 if rng.integers(2) else -1)
        if mode == 1:      # nearly equal values
            b[i] = a[i]
            b[i, 1 + int(rng.integers(nl))] ^= np.uint64(1) << np.uint64(int(rng.integers(64)))
            eb, sb = ea, sa
        elif mode == 2:    # x+1 000.. vs x fff..
            a[i, 1:nl] = 0
            b[i, 1:nl] = np.uint64(0xFFFFFFFFFFFFFFFF)
            a[i, nl] = 5
            b[i, nl] = 4
            eb, sb = ea, sa
        elif mode == 3:    # 1 000.. (e+1) vs fff.. (e)
            a[i, 1:nl] = 0
            a[i, nl] = 1
            b[i, 1:nl + 1] = np.uint64(0xFFFFFFFFFFFFFFFF)
            b[i, 1] = rng.integers(1 << 62)
            ea, sb = eb + 1, sa
        elif mode == 4:    # short operands (trailing zero limbs)
            a[i, 1:nl - 1] = 0
            b[i, 1:nl - 2] = 0
        elif mode == 5:    # exact zero
            b[i] = 0
            a[i, 0] = np.uint64((ea & 0xFFFFFFFF) | ((sa & 0xFFFFFFFF) << 32))
            continue
        if b[i, nl] == 0:
            b[i, nl] = 1
        a[i, 0] = np.uint64((ea & 0xFFFFFFFF) | ((sa & 0xFFFFFFFF) << 32))
        b[i, 0] = np.uint64((eb & 0xFFFFFFFF) | ((sb & 0xFFFFFFFF) << 32))
    return a, b


@pytest.mark.parametrize("prec", [128, 448, 768, 1536])
def test_device_scalar_arithmetic_matches_libgmp(prec):
    ctx = sdpb_b200.SchurContext(prec, [(1, 2)], 1)
    a, b = _adversarial_operands(prec, 4096, 11)
    for op, k in [(0, 0), (1, 0), (2, 0), (3, 0), (5, prec), (5, 77), (6, prec), (6, 2 * prec), (6, 13), (7, 0)]:
        got = ctx.scalar_op(op, a, b, k)
        want = ol.scalar_op(prec, op, a, b, k)
        bad = np.argwhere((got != want).any(axis=1))
        assert len(bad) == 0, f"op {op} k {k}: {len(bad)} mismatches, first index {bad[0]}"
    # sqrt needs a >= 0
    a[:, 0] = (a[:, 0] & np.uint64(0xFFFFFFFF)) | (np.uint64(1) << np.uint64(32))
    got = ctx.scalar_op(4, a, b)
    want = ol.scalar_op(prec, 4, a, b)
    assert np.array_equal(got, want)
    ctx.close()


@pytest.mark.parametrize("prec", [128, 448, 768, 960, 1536])
def test_cooperative_pivot_matches_libgmp(prec):
    """coop.cuh: sqrt and pivot reciprocal by one warp (ops 8, 9 of the scalar hook) --
    sqrt against libgmp's mpf_sqrt, the reciprocal words against the single-thread routine,
    whose own fast path must not have fallen back."""
    ctx = sdpb_b200.SchurContext(prec, [(1, 2)], 1)
    a, _ = _adversarial_operands(prec, 2048, 23)
    a[:, 0] = (a[:, 0] & np.uint64(0xFFFFFFFF)) | (np.uint64(1) << np.uint64(32))  # a > 0
    nl = (prec + 63) // 64 + 2
    a[a[:, nl] == 0, nl] = 1
    got = ctx.scalar_op(8, a, a)
    want = ol.scalar_op(prec, 4, a, a)
    bad = np.argwhere((got != want).any(axis=1))
    assert len(bad) == 0, f"cooperative sqrt: {len(bad)} mismatches, first index {bad[0]}"
    got = ctx.scalar_op(9, a, a)
    assert int((got[:, 1] & np.uint64(0xFFFFFFFF)).max()) == 0, "cooperative reciprocal differs"
    ctx.close()


CASES = [
    # prec, [(m, n)...], N
    (128, [(1, 4), (2, 3), (1, 1)], 3),
    (448, [(1, 6), (1, 7), (2, 5)], 5),
    (768, [(1, 5), (2, 4), (3, 2)], 4),
    (768, [(1, 24), (1, 25), (1, 31)], 20),     # C1-like blocks (J=11,N=20 fixture shapes)
    (664, [(1, 5)], 1),                          # the `1d` fixture's precision (not a multiple of 64)
    (960, [(2, 6), (1, 9)], 7),
]


@pytest.mark.parametrize("prec,shapes,N", CASES)
def test_schur_step_bit_exact(prec, shapes, N):
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=3)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    got = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k, got[k], want[k])
    # second step on the same context must reproduce itself (state is reset), here with
    # every kernel on one stream instead of the concurrent schedule
    ctx.set_concurrency(0)
    again = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k + " (2nd step, single stream)", again[k], want[k])
    ctx.close()


SOLVE_CASES = CASES + [
    (768, [(1, 40), (2, 40), (1, 33), (2, 17)], 37),   # several 16-row tiles, more rows than a warp
    (1536, [(1, 9), (2, 4)], 6),
]


@pytest.mark.parametrize("prec,shapes,N", SOLVE_CASES)
def test_schur_solve_bit_exact(prec, shapes, N):
    """solve_schur_complement_equation.cxx:16-79 on the device-resident L_j, L_j^-1 B_j, chol(Q)
     against the oracle's restatement, every limb of dx and dy; twice with
    different right-hand sides (the predictor and the corrector of one iteration)."""
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=3)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    sdp.run_step(ctx)
    for seed in (7, 8):
        want_dx, want_dy = sdp.solve_rhs(seed)
        ref.solve_schur_complement_equation(want_dx, want_dy)
        dx, dy = sdp.solve_rhs(seed)
        ctx.solve_schur_complement_equation(dx, dy)
        ol.assert_same("dy", dy, want_dy)
        ol.assert_same("dx", dx, want_dx)
    assert ctx.last_solve_ms() > 0
    # zero right-hand side: exact zeros stay exact zeros
    dx, dy = ctx.alloc_solve_vectors()
    ctx.solve_schur_complement_equation(dx, dy)
    assert not any(a.any() for a in dx) and not dy.any()
    ctx.close()


def test_schur_solve_wide_Q_and_global_memory_path(monkeypatch):
    """N = 530 > 512 threads: several rows per thread in the substitution with chol(Q) (the c4 shape
    class, N = 1000); then the same solve with the unknowns in global instead of shared memory
    (what systems beyond 200 KB of unknowns use)."""
    prec, shapes, N = 448, [(2, 40)] * 5, 530
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=9)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    sdp.run_step(ref)
    want_dx, want_dy = sdp.solve_rhs()
    ref.solve_schur_complement_equation(want_dx, want_dy)
    for smem_max in (None, "0"):
        if smem_max is not None:
            monkeypatch.setenv("SDPB_B200_SOLVE_SMEM_MAX", smem_max)
        ctx = sdpb_b200.SchurContext(prec, shapes, N)
        sdp.upload(ctx)
        ctx.upload_XY(sdp.X, sdp.Y)
        ctx.schur_step_resident()
        dx, dy = sdp.solve_rhs()
        ctx.solve_schur_complement_equation(dx, dy)
        ol.assert_same(f"dy (smem_max={smem_max})", dy, want_dy)
        ol.assert_same(f"dx (smem_max={smem_max})", dx, want_dx)
        ctx.close()


def test_schur_solve_after_resident_step_and_state():
    prec, shapes, N = 768, [(1, 12), (2, 7)], 9
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=4)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    dx, dy = sdp.solve_rhs()
    with pytest.raises(sdpb_b200.SdpbB200Error) as ei:
        ctx.solve_schur_complement_equation(dx, dy)   # no factors yet
    assert ei.value.code == 5
    ctx.upload_XY(sdp.X, sdp.Y)
    ctx.schur_step_resident()
    ctx.solve_schur_complement_equation(dx, dy)
    want_dx, want_dy = sdp.solve_rhs()
    ref.solve_schur_complement_equation(want_dx, want_dy)
    ol.assert_same("dy", dy, want_dy)
    ol.assert_same("dx", dx, want_dx)
    ctx.close()


@pytest.mark.parametrize("prec,shapes,N", [CASES[0], CASES[2], CASES[3], (1536, [(1, 9), (2, 4)], 6),
                                           (768, [(2, 40), (1, 40), (1, 33), (2, 17)], 3)])
def test_scale_multiply_add_bit_exact(prec, shapes, N):
    """scale_multiply_add.cxx:4-16 for the reference's three (alpha, beta):
    -X Y (step.cxx:137), (1, 0) and (-1, 1) (compute_search_direction.cxx:28,60)."""
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=6)
    ref = ol.OracleContext(prec, shapes, N)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    for alpha, beta in ((-1, 0), (1, 0), (-1, 1)):
        want = [a.copy() for a in sdp.X]
        ref.scale_multiply_add(alpha, sdp.X, sdp.Y, beta, want)
        got = [a.copy() for a in sdp.X]
        ctx.scale_multiply_add(alpha, sdp.X, sdp.Y, beta, got)
        ol.assert_same(f"C (alpha={alpha}, beta={beta})", got, want)
    with pytest.raises(sdpb_b200.SdpbB200Error) as ei:
        ctx.scale_multiply_add(3, sdp.X, sdp.Y, 0, got)
    assert ei.value.code == 1
    ctx.close()


@pytest.mark.parametrize("groups", ["3", "size"])
def test_block_groups_on_side_streams_bit_exact(monkeypatch, groups):
    """The S chain cut into groups of blocks on separate streams (SDPB_B200_GROUPS) must not change
    a bit: interleaved groups ("3"), and the split by size class that is the default for batches
    with two classes of blocks ("size" forces it at this small block count: the 33- and 27-row
    blocks against the rest), also on a second step and on a single stream."""
    prec, shapes, N = 768, [(1, 9), (2, 5), (1, 17), (1, 4), (2, 11), (1, 12), (1, 3), (2, 9)], 9
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=8)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    monkeypatch.setenv("SDPB_B200_GROUPS", groups)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    got = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k, got[k], want[k])
    ctx.set_concurrency(0)
    again = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k + " (2nd step, single stream)", again[k], want[k])
    want_dx, want_dy = sdp.solve_rhs()
    ref.solve_schur_complement_equation(want_dx, want_dy)
    dx, dy = sdp.solve_rhs()
    ctx.solve_schur_complement_equation(dx, dy)
    ol.assert_same("dy", dy, want_dy)
    ol.assert_same("dx", dx, want_dx)
    ctx.close()


@pytest.mark.parametrize("below", ["0", "1000000000"])
def test_trsm_diagonal_solve_forms_bit_exact(monkeypatch, below):
    """The diagonal solves of L^-1 B and L_X^-1 V have two schedules -- one thread per column
    (trsm_diag_level, full-size batches) and 16 x 16 threads per tile (trsm_diag_tile, batches too
    small to fill the SMs) -- chosen by the column count of a level; SDPB_B200_TRSM_TILE_BELOW
    forces either.  Both must reproduce the oracle bit for bit (ragged last tiles, m = 2 blocks,
    a column count that is not a multiple of 16)."""
    prec, shapes, N = 768, [(1, 40), (2, 17), (1, 9), (2, 6), (1, 33)], 21
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=12)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    monkeypatch.setenv("SDPB_B200_TRSM_TILE_BELOW", below)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    got = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k, got[k], want[k])
    ctx.close()


@pytest.mark.parametrize("fused", ["0", "1"])
def test_cholesky_Q_pipelined_and_two_kernel_forms_bit_exact(monkeypatch, fused):
    """Cholesky(Q) runs its diagonal tile and the panel below it in one launch, the panel one column
    behind (potrf_diag_panel_rl); SDPB_B200_POTRF_FUSED=0 keeps the two kernels.  N = 53 gives four
    levels with a ragged last tile; both forms must reproduce the oracle, on a second step too."""
    prec, shapes, N = 768, [(1, 30), (2, 9), (1, 17)], 53
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=21)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    monkeypatch.setenv("SDPB_B200_POTRF_FUSED", fused)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    for rep in range(2):
        got = sdp.run_step(ctx)
        for k in KEYS:
            ol.assert_same(k + f" (step {rep})", got[k], want[k])
    ctx.close()


def test_rank_deficient_Q_ends_like_the_oracle():
    """Fewer stacked rows than columns (K = 11 < N = 20): Q = P^T P is singular, the pivots from
    column 12 on are rounding noise, and whether -- and where -- Cholesky(Q) meets a non-positive
    one is decided by the arithmetic.  The CUDA path must end exactly like the oracle: the same
    error (a failing pivot inside the first diagonal tile also wakes the waiting panel CTA of the
    pipelined kernel) or the same bits."""
    prec, shapes, N = 768, [(1, 5), (1, 6)], 20
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=4)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want, want_err = None, None
    try:
        want = sdp.run_step(ref)
    except ol.OracleError as e:
        want_err = e
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    if want_err is None:
        got = sdp.run_step(ctx)
        for k in KEYS:
            ol.assert_same(k, got[k], want[k])
    else:
        with pytest.raises(sdpb_b200.SdpbB200Error) as ei:
            sdp.run_step(ctx)
        assert ei.value.code == want_err.code
        # the reference's text (initialize_schur_complement_solver.cxx:100-103); the CUDA path appends the pivot
        assert ei.value.message.startswith(want_err.message)
    ctx.close()


def test_separate_calls_match_fused_step():
    prec, shapes, N = 256, [(1, 6), (2, 3)], 4
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=5)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    fused = sdp.run_step(ctx)
    Xc, Yc = ctx.alloc_psd_blocks(), ctx.alloc_psd_blocks()
    AX, AY = ctx.alloc_pairing_blocks(), ctx.alloc_pairing_blocks()
    L, P, Q = ctx.alloc_schur_outputs()
    ctx.cholesky_decomposition(0, sdp.X, Xc)
    ctx.cholesky_decomposition(1, sdp.Y, Yc)
    ctx.compute_bilinear_pairings(sdp.Y, AX, AY)
    bt = np.zeros(len(shapes), dtype=np.int32)
    ctx.initialize_schur_complement_solver(L, P, Q, bt)
    for k, v in zip(KEYS, [Xc, Yc, AX, AY, L, P, Q]):
        ol.assert_same(k, v, fused[k])
    t = ctx.last_timings_ms()
    assert t[8] > 0
    ctx.close()


def test_non_pd_inputs_raise_reference_style_errors():
    prec, shapes, N = 128, [(1, 4), (1, 3)], 2
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=2)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    good = sdp.X[3]
    sdp.X[3] = ol.scale_matrix(prec, good, -1.0)
    with pytest.raises(sdpb_b200.SdpbB200Error) as ei:
        sdp.run_step(ctx)
    assert ei.value.code == 3
    assert "Block_Diagonal_Matrix X, block index = 1, parity = 1" in str(ei.value)
    sdp.X[3] = good
    sdp.Y[0] = ol.scale_matrix(prec, sdp.Y[0], -1.0)
    with pytest.raises(sdpb_b200.SdpbB200Error) as ei:
        sdp.run_step(ctx)
    assert "Block_Diagonal_Matrix Y, block index = 0, parity = 0" in str(ei.value)
    ctx.close()


def test_calls_out_of_order_are_rejected():
    ctx = sdpb_b200.SchurContext(128, [(1, 3)], 2)
    with pytest.raises(sdpb_b200.SdpbB200Error) as ei:
        ctx.compute_bilinear_pairings(ctx.alloc_psd_blocks())
    assert ei.value.code == 5
    ctx.close()


def test_c3_sample_full_width_bit_exact():
    """The bench workload's own shapes at its full width (N = 300 columns, P_j = 40 / 120, 768 bits;
    24 of the 600 blocks so that the CPU oracle finishes in seconds): every output bit for bit,
    through the host-buffer pipeline (sdpb_b200_schur_step) and through the split calls."""
    from sdpb_b200.synthetic import WORKLOADS
    prec, shapes, N = WORKLOADS["c3-sample"]
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=1)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    got = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k, got[k], want[k])
    want_dx, want_dy = sdp.solve_rhs()
    ref.solve_schur_complement_equation(want_dx, want_dy)
    dx, dy = sdp.solve_rhs()
    ctx.solve_schur_complement_equation(dx, dy)
    ol.assert_same("dy", dy, want_dy)
    ol.assert_same("dx", dx, want_dx)
    ctx.close()


def test_full_c3_properties():
    """BASELINE config c3 at full size (J = 600, N = 300, 768 bits) through size-independent
    properties: (1) idempotence -- two steps from the same X, Y give identical bytes, in the
    concurrent and in the single-stream schedule; (2) the per-block outputs of blocks that also
    occur in a small SDP (same seed => same data) do not depend on the other 576 blocks up to the
    stage where the column norms couple them: X/Y Cholesky factors, pairings and L_j are compared
    bit for bit with the oracle run on the 24-block sub-problem."""
    from sdpb_b200.synthetic import WORKLOADS
    prec, shapes, N = WORKLOADS["c3"]
    pick = list(range(0, 6)) + list(range(150, 168))  # 6 blocks m=2, 18 blocks m=1
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=1)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    got = sdp.run_step(ctx)
    ctx.set_concurrency(0)
    again = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k + " (idempotence / schedule)", again[k], got[k])
    ctx.close()
    sub_shapes = [shapes[j] for j in pick]
    sub = ol.SyntheticSDP(prec, sub_shapes, N, seed=1, block_ids=pick)
    for a, j in zip(sub.B, pick):
        assert np.array_equal(a, sdp.B[j])
    ref = ol.OracleContext(prec, sub_shapes, N)
    sub.upload(ref)
    want = sub.run_step(ref)
    for k in ("X_chol", "Y_chol", "A_X_inv", "A_Y"):
        ol.assert_same(k, [got[k][2 * j + p] for j in pick for p in (0, 1)], want[k])
    ol.assert_same("L", [got["L"][j] for j in pick], want["L"])


def test_cholesky_diagonals_are_the_factor_diagonals():
    """sdpb_b200_cholesky_diagonals (what update_cond_numbers reads, step.cxx:187-189): the stacked
    diagonals of X/Y factors, L_j and chol(Q) equal those of the factors the oracle computes."""
    prec, shapes, N = 768, [(1, 7), (2, 5), (1, 9), (2, 3)], 9
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=4)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    ctx.schur_step(sdp.X, sdp.Y)  # nothing but the status crosses PCIe
    ew = ctx.ew
    nxy = sum(s.psd_size(p) for s in ctx.shapes for p in (0, 1))
    K = sum(s.schur_size for s in ctx.shapes)
    Xd, Yd = np.zeros((nxy, ew), np.uint64), np.zeros((nxy, ew), np.uint64)
    Sd, Qd = np.zeros((K, ew), np.uint64), np.zeros((N, ew), np.uint64)
    ctx.cholesky_diagonals(Xd, Yd, Sd, Qd)

    def diags(blocks):
        return np.concatenate([np.stack([b[i, i] for i in range(b.shape[0])]) for b in blocks if b.shape[0]])

    assert np.array_equal(Xd, diags(want["X_chol"]))
    assert np.array_equal(Yd, diags(want["Y_chol"]))
    assert np.array_equal(Sd, diags(want["L"]))
    assert np.array_equal(Qd, diags([want["Q"]]))
    ctx.close()


def test_any_precision_kernels_are_built_on_demand():
    """--precision is free in the reference (Environment.cxx:29-36).  320 bits (7 stored limbs) is
    not linked into libsdpb_b200.so: the library builds libsdpb_b200_nl7.so with nvcc on first use
    (a minute at this size), loads it, and the step is bit-exact like any other precision."""
    prec, shapes, N = 320, [(1, 6), (2, 4), (1, 9)], 5
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=6)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    got = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k, got[k], want[k])
    a, b = _adversarial_operands(prec, 1024, 5)
    for op in (0, 1, 2, 3):
        assert np.array_equal(ctx.scalar_op(op, a, b), ol.scalar_op(prec, op, a, b))
    ctx.close()


@pytest.mark.parametrize("name,cut", [("c2", 12), ("c4-sample", None)])
def test_named_config_shapes_bit_exact(name, cut):
    """BASELINE configs C2 (448 bits, N = 60, P_j = 30) and C4 (960 bits, N = 1000) at their own
    block shapes and full width; a dozen / two dozen blocks so that the oracle finishes in seconds
    (C4's bands then have more than N rows, as Q must be positive definite)."""
    from sdpb_b200.synthetic import WORKLOADS
    prec, shapes, N = WORKLOADS[name]
    shapes = shapes[:cut] if cut else shapes
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=2)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    want = sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    got = sdp.run_step(ctx)
    for k in KEYS:
        ol.assert_same(k, got[k], want[k])
    want_dx, want_dy = sdp.solve_rhs()
    ref.solve_schur_complement_equation(want_dx, want_dy)
    dx, dy = sdp.solve_rhs()
    ctx.solve_schur_complement_equation(dx, dy)
    ol.assert_same("dy", dy, want_dy)
    ol.assert_same("dx", dx, want_dx)
    ctx.close()


def test_schur_solve_takes_the_reduced_residue_of_the_references_layout():
    """The reference stores r_y as one share per block (dy.blocks[j] = -B_j^T x_j, plus b on global
    block 0; compute_primal_residues_and_error_p_b_Bx.cxx:24-34) and sums the shares inside
    solve_schur_complement_equation.cxx:26-60.  The C-ABI takes the REDUCED vector (INTEGRATION.md's
    shim adds the shares and all-reduces them): build the per-block shares, reduce them as the shim
    does, and check the device solve against the oracle fed the same reduced r_y -- and that feeding
    only the first block's share, the mistake an integrator could make, gives a different answer."""
    prec, shapes, N = 768, [(1, 7), (2, 5), (1, 9)], 6
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=8)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    sdp.run_step(ctx)
    ew = ctx.ew
    # per-block shares of r_y: random vectors standing in for -B_j^T x_j (+ b on block 0)
    shares = [ol.random_matrix(prec, N, 1, 900 + j) for j in range(len(shapes))]
    r_y = shares[0].reshape(N, ew).copy()
    for sh in shares[1:]:
        r_y = ol.scalar_op(prec, 1, r_y, sh.reshape(N, ew))   # mpf_add, block order
    r_y = r_y.reshape(1, N, ew)
    want_dx, _ = sdp.solve_rhs()
    want_dy = r_y.copy()
    ref.solve_schur_complement_equation(want_dx, want_dy)
    dx, _ = sdp.solve_rhs()
    dy = r_y.copy()
    ctx.solve_schur_complement_equation(dx, dy)
    ol.assert_same("dy", dy, want_dy)
    ol.assert_same("dx", dx, want_dx)
    dx1, _ = sdp.solve_rhs()
    dy1 = shares[0].copy()
    ctx.solve_schur_complement_equation(dx1, dy1)
    assert not np.array_equal(dy1, dy)
    ctx.close()


def _random_residues(prec, ctx, N, seed):
    pr = [ol.random_matrix(prec, s.psd_size(p), s.psd_size(p), seed + 17 * b + p)
          for b, s in enumerate(ctx.shapes) for p in (0, 1)]
    dr = [ol.random_matrix(prec, s.schur_size, 1, seed + 1000 + j) for j, s in enumerate(ctx.shapes)]
    return pr, dr, ol.random_matrix(prec, N, 1, seed + 5000)


DIRECTION_CASES = [
    (768, [(1, 7), (2, 5), (1, 9), (2, 3)], 9),
    (448, [(1, 6), (3, 3), (2, 4)], 5),
    (1536, [(1, 9), (2, 4)], 6),
    (664, [(1, 5), (1, 1)], 2),          # not a multiple of 64; a block whose odd parity is empty
    (768, [(2, 40), (1, 40), (1, 33)], 37),
]


@pytest.mark.parametrize("prec,shapes,N", DIRECTION_CASES)
def test_search_direction_bit_exact(prec, shapes, N):
    """Rows N2: compute_search_direction.cxx:44-90 (R, Z, cholesky_solve, symmetrize, compute_schur_RHS,
    the Schur solve, constraint_matrix_weighted_sum, dY) and the per-block reductions of step()
    (traces of -XY, R error, Frobenius product) on the device-resident objects, against the host
    restatement of csrc/host/direction.hpp: predictor, then corrector (which reads the predictor's
    dX, dY), every output byte for byte."""
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=9)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    ctx.schur_step(sdp.X, sdp.Y)
    ol.assert_same("traces", ctx.direction_begin(), ref.direction_begin())
    mu = ol.from_decimal(prec, "0.37251")
    ol.assert_same("R errors", ctx.direction_R_errors(mu), ref.direction_R_errors(mu))
    pr, dr, prp = _random_residues(prec, ctx, N, 77)
    ctx.direction_set_residues(pr, dr, prp)
    ref.direction_set_residues(pr, dr, prp)
    for phase, beta_mu in ((0, "0.1117"), (1, "0.0433")):
        bm = ol.from_decimal(prec, beta_mu)
        ctx.compute_search_direction(bm, phase)
        ref.compute_search_direction(bm, phase)
        got, want = ctx.direction_get(), ref.direction_get()
        for name, g, w in zip(("dx", "dX", "dy", "dY"), got, want):
            ol.assert_same(f"{name} (phase {phase})", g, w)
        ol.assert_same(f"Frobenius products (phase {phase})", ctx.direction_frobenius(), ref.direction_frobenius())
        # row N3: step_length.cxx:27-46 on the resident factors and direction -- congruence with chol(X)
        # / chol(Y), tridiagonalisation, Laguerre's iteration -- against csrc/host/step_length.hpp
        for which in (0, 1):
            ol.assert_same(f"min eigenvalues of L^-1 d{'XY'[which]} L^-T (phase {phase})",
                           ctx.step_length(which), ref.step_length(which))
    ctx.close()


def test_search_direction_after_separate_calls_and_state_errors():
    """The direction also runs on the X, Y of cholesky_decomposition(X) + compute_bilinear_pairings(Y)
    + initialize_schur_complement_solver (the reference's call sequence), and out-of-order calls are
    rejected with rc 5."""
    from sdpb_b200.capi import SdpbB200Error
    prec, shapes, N = 768, [(1, 7), (2, 5), (1, 9)], 8
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=12)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    with pytest.raises(SdpbB200Error) as ei:
        ctx.direction_begin()
    assert ei.value.code == 5
    ctx.cholesky_decomposition(0, sdp.X)
    ctx.cholesky_decomposition(1, sdp.Y)
    ctx.compute_bilinear_pairings(sdp.Y)
    ctx.initialize_schur_complement_solver()
    bm = ol.from_decimal(prec, "0.25")
    with pytest.raises(SdpbB200Error) as ei:
        ctx.compute_search_direction(bm, 0)      # before direction_begin
    assert ei.value.code == 5
    with pytest.raises(SdpbB200Error) as ei:
        ctx.step_length(0)                       # row N3 needs a direction (computed or put)
    assert ei.value.code == 5
    ol.assert_same("traces", ctx.direction_begin(), ref.direction_begin())
    with pytest.raises(SdpbB200Error) as ei:
        ctx.compute_search_direction(bm, 0)      # before the residues
    assert ei.value.code == 5
    pr, dr, prp = _random_residues(prec, ctx, N, 5)
    ctx.direction_set_residues(pr, dr, prp)
    ref.direction_set_residues(pr, dr, prp)
    with pytest.raises(SdpbB200Error) as ei:
        ctx.compute_search_direction(bm, 1)      # corrector before predictor
    assert ei.value.code == 5
    ctx.compute_search_direction(bm, 0)
    ref.compute_search_direction(bm, 0)
    for name, g, w in zip(("dx", "dX", "dy", "dY"), ctx.direction_get(), ref.direction_get()):
        ol.assert_same(name, g, w)
    ctx.close()


def test_step_length_degenerate_spectra_bit_exact():
    """Row N3 on directions put in by the caller (sdpb_b200_direction_put): an exact zero block, a
    multiple of X (L^-1 dX L^-T = c I up to rounding: the spectrum Laguerre's iteration is slowest
    on), a diagonal block, a tiny and a huge block, a block that is not symmetric (the lower
    triangle is the one El::HermitianEig(LOWER) reads), 1 x 1 and 2 x 2 blocks and several warps of
    rows -- every eigenvalue byte for byte against csrc/host/step_length.hpp."""
    prec, shapes, N = 768, [(1, 2), (1, 3), (2, 40), (1, 40), (1, 33), (1, 7), (2, 5), (1, 9)], 11
    ew = elem_words(prec)
    sdp = ol.SyntheticSDP(prec, shapes, N, seed=21)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    sdp.run_step(ref)
    ctx = sdpb_b200.SchurContext(prec, shapes, N)
    sdp.upload(ctx)
    ctx.schur_step(sdp.X, sdp.Y)

    def crafted(M, seed):
        out = []
        for b, blk in enumerate(M):
            s = blk.shape[0]
            kind = b % 6
            if s == 0 or kind == 0:
                d = np.zeros_like(blk)                                   # exact zero
            elif kind == 1:
                d = ol.scale_matrix(prec, blk, -2.5)                     # c * M
            elif kind == 2:
                d = np.zeros_like(blk)                                   # diagonal, distinct entries
                for i in range(s):
                    d[i, i] = ol.from_decimal(prec, str((-1) ** i * (i + 1) * 0.37))
            elif kind == 3:
                r = ol.random_matrix(prec, s, s, seed + b)               # not symmetric
                d = ol.scale_matrix(prec, r, 1e-40)
            elif kind == 4:
                r = ol.random_matrix(prec, s, s, seed + b)
                rt = np.ascontiguousarray(np.transpose(r, (1, 0, 2)))
                d = ol.scale_matrix(prec, ol.scalar_op(prec, 1, r.reshape(-1, ew), rt.reshape(-1, ew)).reshape(s, s, ew), 1e30)
            else:
                d = ol.scale_matrix(prec, blk, 1.0)                      # + M itself: min eigenvalue 1
            out.append(np.ascontiguousarray(d))
        return out

    for seed in (3, 4):
        dX, dY = crafted(sdp.X, 100 * seed), crafted(sdp.Y, 100 * seed + 50)
        ctx.direction_put(dX, dY)
        ref.direction_put(dX, dY)
        for which in (0, 1):
            got, want = ctx.step_length(which), ref.step_length(which)
            ol.assert_same(f"min eigenvalues, which={which}", got, want)
    ctx.close()
