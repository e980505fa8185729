"""One rank of the sharded Schur step; launched by tests/test_sharded.py as
  python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 ... tests/sharded_worker.py MODE

MODE = oracle : gloo, CPU.  The blocks are split with sdpb_b200.partition; each rank runs the oracle's
                staged model on its share and the two exchanges of DESIGN.md §7 go through
                torch.distributed (all_gather).  Checks the sharding logic and exchange semantics.
MODE = b200q  : as b200, with Cholesky(Q) forced onto the panel-distributed path (N = 41).
MODE = b200   : nccl, one GPU per rank.  Each rank creates a SchurContext with its blocks, joins the
                library's NCCL communicator (sdpb_b200_comm_init) and runs the collective step.
Either way every rank compares its outputs bit for bit with the UNSHARDED oracle on the whole SDP."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib as ol  # noqa: E402
from sdpb_b200.partition import partition_blocks  # noqa: E402

PREC, N = 768, 7
SHAPES = [(1, 6), (2, 4), (1, 9), (1, 5), (2, 3), (1, 8), (1, 4)]


def main():
    global N
    mode = sys.argv[1]
    if mode == "b200q":
        # Cholesky(Q) by broadcast panels (block columns dealt over the ranks): force the
        # distributed path at a size with three block columns, two of them on rank 0
        os.environ["SDPB_B200_QDIST_MIN_N"] = "1"
        N = 41
        mode = "b200"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "b200":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    owned = partition_blocks(SHAPES, N, world)
    assert sorted(sum(owned, [])) == list(range(len(SHAPES)))
    mine = owned[rank]
    shapes = [SHAPES[j] for j in mine]

    # the unsharded answer
    full = ol.SyntheticSDP(PREC, SHAPES, N, seed=5)
    ref = ol.OracleContext(PREC, SHAPES, N)
    full.upload(ref)
    want = full.run_step(ref)
    # ... and the Schur solve (solve_schur_complement_equation) on its factors
    want_dx, want_dy = full.solve_rhs()
    ref.solve_schur_complement_equation(want_dx, want_dy)

    sdp = ol.SyntheticSDP(PREC, shapes, N, seed=5, block_ids=mine)
    for a, j in zip(sdp.B, mine):
        assert np.array_equal(a, full.B[j])
    if mode == "oracle":
        ctx = ol.OracleContext(PREC, shapes, N)
        sdp.upload(ctx)
        part = ctx.shard_stage1(sdp.X, sdp.Y)
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, part))
        part_global = np.zeros((len(SHAPES), N, ctx.ew), dtype=np.uint64)
        for ids, p in gathered:
            for k, j in enumerate(ids):
                part_global[j] = p[k]
        q = ctx.shard_stage2(part_global)
        qs = [torch.zeros(q.shape, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(qs, torch.from_numpy(q.view(np.int64)))
        L, P, Q = ctx.shard_stage3([t.numpy().view(np.uint64) for t in qs])
        # the Schur solve, staged: local forward part, partial rows gathered in global block order
        dx, dy = sdp.solve_rhs()
        spart = ctx.shard_solve_stage1(dx)
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, spart))
        spart_global = np.zeros((len(SHAPES), N, ctx.ew), dtype=np.uint64)
        for ids, p in gathered:
            for k, j in enumerate(ids):
                spart_global[j] = p[k]
        ctx.shard_solve_stage2(spart_global, dx, dy)
    else:
        import sdpb_b200
        ctx = sdpb_b200.SchurContext(PREC, shapes, N, device=local)
        ctx.comm_init_from_torch(dist, rank, world, len(SHAPES), mine)
        sdp.upload(ctx)
        got = sdp.run_step(ctx)
        L, P, Q = got["L"], got["P"], got["Q"]
        for k in ("X_chol", "Y_chol", "A_X_inv", "A_Y"):
            ol.assert_same(k, got[k], [want[k][2 * j + p] for j in mine for p in (0, 1)])
        # a second step on the same communicator (buffers are reused)
        got2 = sdp.run_step(ctx)
        ol.assert_same("Q(second step)", got2["Q"], want["Q"])
        dx, dy = sdp.solve_rhs()
        ctx.solve_schur_complement_equation(dx, dy)
    ol.assert_same("Q", Q, want["Q"])
    ol.assert_same("L", L, [want["L"][j] for j in mine])
    ol.assert_same("P", P, [want["P"][j] for j in mine])
    ol.assert_same("dy", dy, want_dy)
    ol.assert_same("dx", dx, [want_dx[j] for j in mine])
    dist.barrier()
    print(f"rank {rank}/{world} mode {sys.argv[1]}: blocks {mine} match the unsharded oracle bit for bit", flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
