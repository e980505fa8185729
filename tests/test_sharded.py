"""N > 1: the SDP blocks sharded over ranks, world_size 2.

  not gpu : gloo on the CPU — the sharding logic and the semantics of the two exchanges
            (norm partials gathered and summed in global block order; exact integer Q' partials
            added) on the oracle's staged model, against the unsharded oracle, bit for bit.
  gpu     : nccl, 2 GPUs — the library's own communicator (sdpb_b200_comm_init) against the
            unsharded oracle, bit for bit.  Skipped on a 1-GPU box.
"""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(mode, world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "sharded_worker.py"), mode]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(world):
        assert f"rank {rank}/{world} mode {mode}" in r.stdout, r.stdout[-2000:]


def test_partition_is_balanced_and_complete():
    from sdpb_b200.partition import block_cost, partition_blocks
    shapes = [(2, 40)] * 150 + [(1, 40)] * 450
    for world in (1, 2, 4, 8):
        owned = partition_blocks(shapes, 300, world)
        assert sorted(sum(owned, [])) == list(range(600))
        assert all(o == sorted(o) for o in owned)
        loads = [sum(block_cost(*shapes[j], 300) for j in o) for o in owned]
        assert max(loads) <= 1.02 * min(loads)
    assert partition_blocks([(1, 3)], 2, 4) == [[0], [], [], []]


def test_sharded_oracle_model_gloo_world2(oracle):
    _launch("oracle", 2)


@pytest.mark.gpu
def test_sharded_b200_nccl_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    _launch("b200", 2)


@pytest.mark.gpu
def test_sharded_b200_panel_distributed_Q_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    _launch("b200q", 2)
