"""Block -> GPU assignment for the sharded Schur step.

The reference bin-packs SDP blocks onto MPI rank groups by a measured cost
(src/sdpb_util/block_mapping/compute_block_grid_mapping.hxx:58-183,
src/sdp_solve/Block_Info/allocate_blocks.cxx:6-75).  Here the cost is the
limb-MAC model of SURVEY.md §8(e) and the packing is plain LPT (largest first
onto the least loaded rank); every rank keeps its blocks in ascending global
order, which is the order the column-norm partials are summed in.
"""


def block_cost(m, n, N):
    P = n * m * (m + 1) // 2
    s0 = m * ((n + 1) // 2)
    s1 = m * n - s0
    mn = m * n
    pair = sum(s * s * s / 3 * 2 + s * s * mn * 1.5 + mn * mn * s * 1.5 for s in (s0, s1))
    return P ** 3 / 3 + P * P * N / 2 + P * N * N / 2 + 8 * P * P + pair


def partition_blocks(shapes, N, world):
    """shapes: [(m, n)] of all J blocks -> list (per rank) of ascending global block indices."""
    order = sorted(range(len(shapes)), key=lambda j: (-block_cost(shapes[j][0], shapes[j][1], N), j))
    load = [0.0] * world
    owned = [[] for _ in range(world)]
    for j in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owned[r].append(j)
        load[r] += block_cost(shapes[j][0], shapes[j][1], N)
    return [sorted(o) for o in owned]
