"""The sharded Schur step on ONE GPU: every rank is a context of this process on device 0, driven by
its own host thread, and the exchanges go through the library's in-process communicator
(sdpb_b200_comm_init_local) instead of NCCL.  Same code path as the multi-GPU run -- per-GLOBAL-block
partial rows summed in global order, exact residue sums, the panel-distributed Cholesky(Q) with its
block-cyclic ownership, the sharded Schur solve -- checked bit for bit against the UNSHARDED oracle.
(tests/test_sharded.py runs the same comparison over NCCL when the box has two GPUs.)"""
import threading

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _run_sharded(prec, shapes, N, world, seed=5, steps=1):
    import sdpb_b200
    from sdpb_b200.partition import partition_blocks
    owned = partition_blocks(shapes, N, world)
    assert sorted(sum(owned, [])) == list(range(len(shapes)))

    full = ol.SyntheticSDP(prec, shapes, N, seed=seed)
    ref = ol.OracleContext(prec, shapes, N)
    full.upload(ref)
    want = full.run_step(ref)
    want_dx, want_dy = full.solve_rhs()
    ref.solve_schur_complement_equation(want_dx, want_dy)

    ctxs, sdps = [], []
    for r in range(world):
        mine = owned[r]
        sdp = ol.SyntheticSDP(prec, [shapes[j] for j in mine], N, seed=seed, block_ids=mine)
        ctx = sdpb_b200.SchurContext(prec, [shapes[j] for j in mine], N, device=0)
        sdp.upload(ctx)
        ctxs.append(ctx)
        sdps.append(sdp)
    sdpb_b200.SchurContext.comm_init_local(ctxs, len(shapes), owned)

    results, errors = [None] * world, [None] * world

    def rank_main(r):
        try:
            for _ in range(steps):
                got = sdps[r].run_step(ctxs[r])
            dx, dy = sdps[r].solve_rhs()
            ctxs[r].solve_schur_complement_equation(dx, dy)
            results[r] = (got, dx, dy)
        except Exception as e:  # noqa: BLE001
            errors[r] = e

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=900)
    try:
        for r in range(world):
            assert errors[r] is None, f"rank {r}: {errors[r]}"
            got, dx, dy = results[r]
            mine = owned[r]
            for k in ("X_chol", "Y_chol", "A_X_inv", "A_Y"):
                ol.assert_same(f"rank{r}.{k}", got[k], [want[k][2 * j + p] for j in mine for p in (0, 1)])
            ol.assert_same(f"rank{r}.L", got["L"], [want["L"][j] for j in mine])
            ol.assert_same(f"rank{r}.P", got["P"], [want["P"][j] for j in mine])
            ol.assert_same(f"rank{r}.Q", got["Q"], want["Q"])
            ol.assert_same(f"rank{r}.dy", dy, want_dy)
            ol.assert_same(f"rank{r}.dx", dx, [want_dx[j] for j in mine])
    finally:
        for c in ctxs:
            c.close()


def test_two_ranks_on_one_gpu_small_shapes():
    _run_sharded(768, [(1, 6), (2, 4), (1, 9), (1, 5), (2, 3), (1, 8), (1, 4)], 7, 2, steps=2)


def test_three_ranks_uneven_shares_and_an_empty_rank():
    # 4 ranks, 3 blocks: one rank owns nothing and still takes part in every exchange
    _run_sharded(448, [(1, 9), (2, 5), (1, 7)], 6, 4)


def test_c3_sample_shapes_two_and_four_ranks():
    from sdpb_b200.synthetic import WORKLOADS
    prec, shapes, N = WORKLOADS["c3-sample"]
    _run_sharded(prec, shapes, N, 2)
    _run_sharded(prec, shapes, N, 4)


def test_panel_distributed_cholesky_Q_from_512_columns_on(monkeypatch):
    # N >= 512 takes the panel-distributed Cholesky(Q) by default (33 block columns dealt over
    # three ranks); the bands must have at least N rows for Q to be positive definite
    monkeypatch.delenv("SDPB_B200_QDIST_MIN_N", raising=False)
    _run_sharded(768, [(2, 40)] * 4 + [(1, 40)] * 3, 528, 3)


def test_panel_distributed_cholesky_Q_forced_small(monkeypatch):
    monkeypatch.setenv("SDPB_B200_QDIST_MIN_N", "1")
    _run_sharded(768, [(1, 6), (2, 4), (1, 9), (1, 5), (2, 3), (1, 8), (1, 7)], 41, 2)


def test_failure_on_one_rank_fails_every_rank():
    """A non-positive-definite X block on rank 1: rank 1 names the (global) block, rank 0 -- whose own
    blocks are fine -- must not return success with a Q built from garbage."""
    import sdpb_b200
    from sdpb_b200.capi import SdpbB200Error
    prec, N = 768, 5
    shapes = [(1, 6), (1, 7), (1, 5), (1, 8)]
    owned = [[0, 2], [1, 3]]
    ctxs, sdps = [], []
    for r in range(2):
        sdp = ol.SyntheticSDP(prec, [shapes[j] for j in owned[r]], N, seed=3, block_ids=owned[r])
        ctx = sdpb_b200.SchurContext(prec, [shapes[j] for j in owned[r]], N, device=0)
        sdp.upload(ctx)
        ctxs.append(ctx)
        sdps.append(sdp)
    # rank 1, local block 1 (global 3), parity 0: make it indefinite by negating it
    bad = sdps[1].X[2].copy()
    hdr = bad[..., 0]
    sign = (hdr >> np.uint64(32)).astype(np.uint32).astype(np.int32)
    hdr[...] = (hdr & np.uint64(0xFFFFFFFF)) | ((-sign).astype(np.uint32).astype(np.uint64) << np.uint64(32))
    sdps[1].X[2] = bad
    sdpb_b200.SchurContext.comm_init_local(ctxs, len(shapes), owned)
    errors = [None, None]

    def rank_main(r):
        try:
            sdps[r].run_step(ctxs[r])
        except SdpbB200Error as e:
            errors[r] = e

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    try:
        assert errors[1] is not None and "block index = 3" in errors[1].message, errors[1]
        assert errors[0] is not None and "peer rank" in errors[0].message, errors[0]
    finally:
        for c in ctxs:
            c.close()


def test_search_direction_sharded_matches_unsharded():
    """Rows N2 under sharding: every rank computes the direction of its own blocks on its resident
    objects; the Schur solve inside is the collective one (dy identical on all ranks).  Against the
    unsharded host restatement, bit for bit."""
    import sdpb_b200
    from sdpb_b200.partition import partition_blocks
    prec, N, world = 768, 9, 2
    shapes = [(1, 6), (2, 4), (1, 9), (1, 5), (2, 3), (1, 8)]
    owned = partition_blocks(shapes, N, world)

    def residues(ids):
        pr = [ol.random_matrix(prec, BlockShape(*shapes[j]).psd_size(p), BlockShape(*shapes[j]).psd_size(p),
                               300 + 17 * j + p) for j in ids for p in (0, 1)]
        dr = [ol.random_matrix(prec, BlockShape(*shapes[j]).schur_size, 1, 900 + j) for j in ids]
        return pr, dr, ol.random_matrix(prec, N, 1, 4242)

    from sdpb_b200.capi import BlockShape
    full = ol.SyntheticSDP(prec, shapes, N, seed=5)
    ref = ol.OracleContext(prec, shapes, N)
    full.upload(ref)
    full.run_step(ref)
    want_tr = ref.direction_begin()
    ref.direction_set_residues(*residues(range(len(shapes))))
    bm = ol.from_decimal(prec, "0.2")
    ref.compute_search_direction(bm, 0)
    ref.compute_search_direction(bm, 1)
    want = ref.direction_get()
    want_min = [ref.step_length(0), ref.step_length(1)]  # row N3: per block-parity, local to the owner

    ctxs, sdps = [], []
    for r in range(world):
        mine = owned[r]
        sdp = ol.SyntheticSDP(prec, [shapes[j] for j in mine], N, seed=5, block_ids=mine)
        ctx = sdpb_b200.SchurContext(prec, [shapes[j] for j in mine], N, device=0)
        sdp.upload(ctx)
        ctxs.append(ctx)
        sdps.append(sdp)
    sdpb_b200.SchurContext.comm_init_local(ctxs, len(shapes), owned)
    results, errors = [None] * world, [None] * world

    def rank_main(r):
        try:
            ctxs[r].schur_step(sdps[r].X, sdps[r].Y)
            tr = ctxs[r].direction_begin()
            ctxs[r].direction_set_residues(*residues(owned[r]))
            ctxs[r].compute_search_direction(bm, 0)
            ctxs[r].compute_search_direction(bm, 1)
            results[r] = (tr, ctxs[r].direction_get(), [ctxs[r].step_length(0), ctxs[r].step_length(1)])
        except Exception as e:  # noqa: BLE001
            errors[r] = e

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    try:
        for r in range(world):
            assert errors[r] is None, f"rank {r}: {errors[r]}"
            tr, (dx, dX, dy, dY), mins = results[r]
            mine = owned[r]
            for which in (0, 1):
                ol.assert_same(f"rank{r}.min eigenvalues {which}", mins[which],
                               np.stack([want_min[which][2 * j + p] for j in mine for p in (0, 1)]))
            ol.assert_same(f"rank{r}.traces", tr, np.stack([want_tr[2 * j + p] for j in mine for p in (0, 1)]))
            ol.assert_same(f"rank{r}.dx", dx, [want[0][j] for j in mine])
            ol.assert_same(f"rank{r}.dX", dX, [want[1][2 * j + p] for j in mine for p in (0, 1)])
            ol.assert_same(f"rank{r}.dy", dy, want[2])
            ol.assert_same(f"rank{r}.dY", dY, [want[3][2 * j + p] for j in mine for p in (0, 1)])
    finally:
        for c in ctxs:
            c.close()
