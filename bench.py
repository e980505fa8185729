#!/usr/bin/env python
"""Benchmark of the Schur-complement step (BASELINE.json: sec per Newton step at
prec = 768 bits; multi-limb Schur kernels' GB/s vs the HBM peak).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One process per GPU (torchrun for N > 1).  A "step" is one pass of the hot path
(cholesky_decomposition(X), (Y), compute_bilinear_pairings,
initialize_schur_complement_solver) over one rank's blocks of a synthetic block
SDP.  Weak scaling: every rank owns one copy of the workload's block list (the
global SDP has N x as many blocks), N fixed columns; the cross-rank exchanges
are the column-norm partials and the exact-syrk residues.

`value`  = seconds per step with X and Y already resident in HBM, timed with
           CUDA events on the library's stream, max over ranks.
`e2e`    = the same step through the reference-facing call sdpb_b200_schur_step
           with HOST buffers: X, Y uploaded from pinned memory and, inside the
           timed region, everything the host side of the solver consumes copied
           back: the Cholesky factors of X and Y and the diagonals of L_j and
           chol(Q) (update_cond_numbers).  L_j, L_j^-1 B_j and chol(Q) themselves
           stay in HBM, where sdpb_b200_solve_schur_complement_equation uses them
           (`e2e_with_two_solves` adds the two solves of an iteration;
           `e2e_all_outputs` is the old contract: every output copied back).
`--impl reference` times the CPU restatement of the reference algorithm
(oracle/, libgmp mpf, OpenMP on every host core the process may run on) on the
FULL block list of the same workload -- under weak scaling the blocks of all N
ranks -- and reports the measured seconds; the reference binary itself cannot be
built in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.samples, self.stop = device, [], False
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in out.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------- algorithmic bytes
def algorithmic_bytes(kernel, prec, shapes, N):
    """Bytes one launch of `kernel` must move, from SURVEY.md §8(d):
    E = 8(L+1)+8 bytes per stored element, L = prec/64+1 limbs."""
    L = (prec + 63) // 64 + 1
    E = 8 * (L + 1) + 8
    tot = 0
    K = sum(s.schur_size for s in shapes)
    for s in shapes:
        P, mn = s.schur_size, s.pairing_size
        for p in (0, 1):
            sp = s.psd_size(p)
            if kernel in ("potrf_X", "potrf_Y"):
                tot += sp * sp * E
            elif kernel == "trsm_LXinv_V":
                tot += (sp * sp // 2 + 2 * sp * mn) * E
            elif kernel == "gemm_A_X_inv":
                tot += (sp * mn + mn * mn) * E
            elif kernel == "gemm_YV":
                tot += (sp * sp + 2 * sp * mn) * E
            elif kernel == "gemm_A_Y":
                tot += (2 * sp * mn + mn * mn) * E
        if kernel == "schur_kernel":
            tot += (4 * mn * mn + P * P) * E
        elif kernel == "potrf_S":
            tot += P * P * E
        elif kernel == "trsm_Linv_B":
            tot += (P * P // 2 + 2 * P * N) * E
        elif kernel in ("norm_partial_kernel",):
            tot += P * N * E
        elif kernel == "normalize_kernel":  # reads P, writes the restored P (+ residues, not counted)
            tot += 2 * P * N * E
    if kernel in ("syrk_mod_kernel", "syrk_imma_kernel"):  # exact integer syrk (SURVEY 8d row "syrk Q'")
        tot = K * N * 8 * L + N * (N + 1) // 2 * 8 * (2 * L + 1)
    elif kernel == "crt_restore_kernel":
        tot = N * (N + 1) // 2 * (8 * (2 * L + 1) + E)
    elif kernel == "potrf_Q":
        tot = N * N * E
    return tot


def limb_macs(kernel, shapes, N):
    """mpf multiply-accumulates of one launch (element updates)."""
    tot = 0
    for s in shapes:
        P, mn = s.schur_size, s.pairing_size
        for p in (0, 1):
            sp = s.psd_size(p)
            if kernel in ("potrf_X", "potrf_Y"):
                tot += sp ** 3 / 3
            elif kernel == "trsm_LXinv_V":
                tot += sp * sp * mn / 2
            elif kernel == "gemm_A_X_inv":
                tot += mn * mn * sp / 2
            elif kernel == "gemm_YV":
                tot += sp * sp * mn
            elif kernel == "gemm_A_Y":
                tot += mn * mn * sp / 2
        if kernel == "schur_kernel":
            tot += 8 * P * P / 2
        elif kernel == "potrf_S":
            tot += P ** 3 / 3
        elif kernel == "trsm_Linv_B":
            tot += P * P * N / 2
    if kernel == "potrf_Q":
        tot = N ** 3 / 3
    return tot


# measured on this pool's B200 (profiles/imad_rate4_r01.jsonl): IMAD.WIDE.U32 with a 64-bit
# accumulate, operands changing every instruction, 28.8 lanes/clk/SM x 148 SMs x 1.965 GHz
IMAD_WIDE_PER_S = 28.77 * 148 * 1.965e9


def imad_per_mac(prec):
    """32x32->64 partial products of one mpf_mul at this precision (mpfw.h short product:
    W = 2(NL-1) operand words, columns from C0 = W-6)."""
    nl = (prec + 63) // 64 + 2
    w = 2 * (nl - 1)
    c0 = max(w - 6, 0)
    return sum(1 for i in range(w) for j in range(w) if i + j >= c0)


def recorded_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of all launches behind `kernel`'s label in one
    step, from the committed `ncu --set full` capture (profiles/traffic_*.json), or None."""
    import glob
    best = None
    import re
    natural = lambda p: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", os.path.basename(p))]
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json")), key=natural):  # the newest capture wins
        try:
            with open(path) as f:
                d = json.load(f)
            if d.get("workload") == workload and d.get("label") == kernel:
                best = float(d["dram_bytes"])
        except Exception:
            pass
    return best


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------ CPU arm
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_all_cores():
    """The oracle library with its OpenMP team set to every core this process may use (torchrun
    exports OMP_NUM_THREADS=1 to its workers, which would silently make the baseline single-core)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    lib = ol.load_oracle()
    lib.oracle_set_num_threads(host_cores())
    return ol, lib.oracle_num_threads()


def cpu_full(workload, steps, warmup, world=1, budget_s=None):
    """Seconds per step of the CPU restatement (oracle/) on the FULL block list of `workload`,
    measured.  Under weak scaling the N-GPU job owns `world` copies of the block list; every stage
    but Cholesky(Q) is linear in the blocks (the exact syrk accumulates over the stacked rows), so
    the CPU processes them as `world` passes over one copy.  The number of timed steps is cut to
    what fits `budget_s` (reported); if not even one full step fits, the bounded block sample is
    timed instead and scaled, and the line says so."""
    ol, cores = oracle_all_cores()
    from sdpb_b200.synthetic import WORKLOADS, SyntheticSDP
    if budget_s is None:
        budget_s = float(os.environ.get("SDPB_B200_CPU_BUDGET_S", "200"))
    prec, shapes, N = WORKLOADS[workload]
    out = {"unit": "s/step", "cores": cores, "kind": "port"}
    # a short sample first: warms the library up and predicts the cost of the full step
    sname = workload + "-sample" if workload + "-sample" in WORKLOADS else None
    est = None
    Ktot = sum(m * (m + 1) // 2 * n for m, n in shapes)
    if sname is None and float(Ktot) * N * N > 2e10:
        # no bounded sample with these shapes exists (N = 4096 needs >= 4096 stacked rows, and the
        # exact syrk + Cholesky(Q) of that alone are CPU-hours): no CPU number rather than a guess
        out.update(value=None, extrapolated=None, steps=0, warmup=0,
                   sample="unavailable: the smallest well-posed sample of this workload is CPU-hours")
        return out
    if sname:
        sprec, sshapes, sN = WORKLOADS[sname]
        sdp = SyntheticSDP(sprec, sshapes, sN, seed=1)
        ref = ol.OracleContext(sprec, sshapes, sN)
        sdp.upload(ref)
        t0 = time.perf_counter()
        ref.schur_step(sdp.X, sdp.Y)
        t_sample = time.perf_counter() - t0
        ms = ref.stage_ms()
        scale = len(shapes) / len(sshapes)
        est = ((ms[0] + ms[1] + ms[2] + ms[3] + ms[5]) * scale * world + ms[7]) / 1e3
        out["sample_estimate"] = {"value": est, "sample_wall_s": t_sample,
                                  "what": f"{len(sshapes)} of {len(shapes)} blocks, per-block stages scaled "
                                          f"x{scale * world:g}, Cholesky(Q) counted once"}
        log(f"[cpu] {sname}: {t_sample:.2f}s -> full-size estimate {est:.1f}s per step on {cores} cores")
        ref.close()
        if est > budget_s:
            out.update(value=est, extrapolated=True, steps=0, warmup=0,
                       sample=f"EXTRAPOLATED from {out['sample_estimate']['what']}: one full step would take "
                              f"{est:.0f}s > budget {budget_s:.0f}s")
            return out
    sdp = SyntheticSDP(prec, shapes, N, seed=1)
    ref = ol.OracleContext(prec, shapes, N)
    sdp.upload(ref)
    sdp.B = None
    walls = []
    t_start = time.perf_counter()
    n_warm = 0
    if warmup > 0 and (est is None or 2 * est * world < budget_s):
        ref.schur_step(sdp.X, sdp.Y)
        n_warm = 1
    for it in range(max(1, steps)):
        t0 = time.perf_counter()
        for _ in range(world):
            ref.schur_step(sdp.X, sdp.Y)
        walls.append(time.perf_counter() - t0)
        log(f"[cpu] full step {it}: {walls[-1]:.2f}s ({world} pass(es) over {len(shapes)} blocks) "
            f"stages(ms) of the last pass {[round(x) for x in ref.stage_ms()]}")
        used = time.perf_counter() - t_start
        if used + walls[-1] > budget_s:
            break
    ref.close()
    out.update(value=float(np.mean(walls)), extrapolated=False, steps=len(walls), warmup=n_warm,
               sample=f"FULL block list, measured: {len(shapes)} blocks x {world} rank(s), N={N}, prec={prec}; "
                      f"{len(walls)} timed step(s) of {float(np.mean(walls)):.1f}s after {n_warm} warm-up(s); "
                      f"exact syrk = direct mpz sum (the faster CPU formulation on this class of host, "
                      f"profiles/cpu_syrk_variants_r01_box.json)")
    return out


def cpu_solve_sample(workload):
    """solve_schur_complement_equation by the CPU restatement on the block sample (after one
    sample step), per-block work scaled to the full block list."""
    ol, _ = oracle_all_cores()
    from sdpb_b200.synthetic import WORKLOADS, SyntheticSDP, solve_rhs
    prec, shapes, N = WORKLOADS[workload]
    sname = workload + "-sample" if workload + "-sample" in WORKLOADS else workload
    sprec, sshapes, sN = WORKLOADS[sname]
    scale = len(shapes) / len(sshapes)
    sdp = SyntheticSDP(sprec, sshapes, sN, seed=1)
    ref = ol.OracleContext(sprec, sshapes, sN)
    sdp.upload(ref)
    ref.schur_step(sdp.X, sdp.Y)
    best = None
    for _ in range(3):
        dx, dy = solve_rhs(sprec, ref.shapes, sN, seed=7)
        t0 = time.perf_counter()
        ref.solve_schur_complement_equation(dx, dy)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": best * scale * 1e3, "unit": "ms/solve", "cores": ol.load_oracle().oracle_num_threads(),
            "kind": "port", "sample": f"{len(sshapes)} of {len(shapes)} blocks, {best * 1e3:.1f} ms, scaled x{scale:g} "
                                      "(the N x N solve with chol(Q), ~10 % of the sample, is scaled too)"}


def cpu_sma_sample(workload):
    """scale_multiply_add(-1, X, Y, 0, C) by the CPU restatement on the block sample, scaled."""
    ol, _ = oracle_all_cores()
    from sdpb_b200.synthetic import WORKLOADS, SyntheticSDP
    prec, shapes, N = WORKLOADS[workload]
    sname = workload + "-sample" if workload + "-sample" in WORKLOADS else workload
    sprec, sshapes, sN = WORKLOADS[sname]
    scale = len(shapes) / len(sshapes)
    sdp = SyntheticSDP(sprec, sshapes, sN, seed=1)
    ref = ol.OracleContext(sprec, sshapes, sN)
    C = [x.copy() for x in sdp.X]
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        ref.scale_multiply_add(-1, sdp.X, sdp.Y, 0, C)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": best * scale * 1e3, "unit": "ms/call", "cores": ol.load_oracle().oracle_num_threads(),
            "kind": "port", "sample": f"{len(sshapes)} of {len(shapes)} blocks, {best * 1e3:.1f} ms, scaled x{scale:g}"}


def bind_to_gpu_numa_node(device):
    """Run this process (and therefore place its pinned staging buffers, first-touch) on the CPUs
    NVML reports as local to `device`.  Eight ranks that all stage through socket 0 share one
    memory controller and the inter-socket link; this is what a launcher's --bind-to does for the
    reference's MPI ranks."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "first": min(cpus), "last": max(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:80]}
    return {"unavailable": "empty affinity mask"}


# --------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("SDPB_B200_WORKLOAD", "c3"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-all-outputs", action="store_true",
                    help="skip the extra e2e pass that copies every output (L_j, L_j^-1 B_j, chol(Q)) back")
    ap.add_argument("--kernels", action="store_true", help="print the per-kernel timeline to stderr")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL writes its version / debug lines to stdout by default; stdout carries the ONE JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # the VERSION banner is a bare printf to stdout
    warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    from sdpb_b200.synthetic import WORKLOADS, SyntheticSDP
    prec, shapes_t, N = WORKLOADS[a.workload]
    cfg = {"workload": f"{a.workload}: shape-synthetic block SDP, J={len(shapes_t)} blocks per GPU "
                       f"({_mix(shapes_t)}), N={N}, prec={prec}",
           "blocks_per_gpu": len(shapes_t), "N": N, "precision_bits": prec,
           "parallelism": f"block-sharded x{world}", "l2": "working set >> 126 MB L2 (B, P, residues: GBs)"}

    if a.impl == "reference":
        if rank != 0:
            return
        cb = cpu_full(a.workload, a.steps, a.warmup, world=max(1, a.gpus))
        line = {"impl": "reference", "metric": "sec_per_newton_step_hot_path", "value": cb["value"], "unit": "s/step",
                "n_gpus": a.gpus, "steps": cb["steps"], "warmup": cb["warmup"],
                "steps_requested": a.steps, "warmup_requested": a.warmup,
                "ms_per_step": cb["value"] * 1e3,
                "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "mpf%d" % prec,
                "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "s/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    numa = bind_to_gpu_numa_node(local)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sdpb_b200
    from sdpb_b200.capi import PinnedPool

    t0 = time.perf_counter()
    sdp = SyntheticSDP(prec, shapes_t, N, seed=1 + rank)
    ctx = sdpb_b200.SchurContext(prec, shapes_t, N, device=local)
    if world > 1:
        # weak scaling: every rank owns its own J blocks of a J*world-block SDP (global index
        # rank*J + j); the column-norm partials and the exact Q' residues cross NVLink (NCCL)
        J = len(shapes_t)
        ctx.comm_init_from_torch(dist, rank, world, J * world, [rank * J + j for j in range(J)])
    sdp.upload(ctx)
    sdp.B = None  # resident in HBM now
    log(f"[rank {rank}] setup {time.perf_counter() - t0:.1f}s")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing (value) ---------------------------------
    ctx.upload_XY(sdp.X, sdp.Y)
    for _ in range(warmup):
        ctx.schur_step_resident()
    launches0 = ctx.kernel_launches()
    dev_ms = []
    barrier()
    with ClockSampler(local) as clk:
        w0 = time.perf_counter()
        for _ in range(a.steps):
            ctx.schur_step_resident()
            dev_ms.append(ctx.last_timings_ms()[8])
        barrier()
        wall = time.perf_counter() - w0
    launches = ctx.kernel_launches() - launches0
    ms_step = float(np.mean(dev_ms))

    # ---- per-kernel timeline: the same step with every kernel on ONE stream in
    # program order (the timed loop above overlaps independent chains on side
    # streams, where a kernel's event-to-event span also contains its neighbours)
    ctx.set_concurrency(0)
    ctx.schur_step_resident()
    ktimes, serial_ms = {}, []
    barrier()
    for _ in range(a.steps):
        ctx.schur_step_resident()
        serial_ms.append(ctx.last_timings_ms()[8])
        for name, ms in ctx.kernel_timings():
            ktimes.setdefault(name, []).append(ms)
    barrier()
    stages = ctx.last_timings_ms()
    serial_step = float(np.mean(serial_ms))
    ctx.set_concurrency(1)

    # ---- end to end through the C-ABI with host buffers -----------------
    pool = PinnedPool()
    # the caller's staging buffers: one pinned slab per block-diagonal object, blocks back to back
    Xh, Yh = pool.slab([x.shape for x in sdp.X]), pool.slab([x.shape for x in sdp.Y])
    for dst, src in zip(Xh + Yh, sdp.X + sdp.Y):
        dst[...] = src
    Xc = pool.slab([x.shape for x in sdp.X])
    Yc = pool.slab([x.shape for x in sdp.X])
    Ktot = sum(s.schur_size for s in ctx.shapes)
    Sdiag, Qdiag = pool.empty((Ktot, ctx.ew)), pool.empty((N, ctx.ew))
    h2d = sum(x.nbytes for x in Xh) + sum(x.nbytes for x in Yh)
    d2h = sum(x.nbytes for x in Xc + Yc) + Sdiag.nbytes + Qdiag.nbytes
    # the block-pointer tables are built once, as a C++ caller's would be (ctypes needs ~5 us per
    # pointer: 6000 pointers per call would put 20-30 ms of Python into the timed region)
    from sdpb_b200.capi import ptr_array
    pX, pY, pXc, pYc = (ptr_array(v) for v in (Xh, Yh, Xc, Yc))

    def e2e_step():
        # L_j, L_j^-1 B_j and chol(Q) stay in HBM for the Schur solves; the host gets the factors of
        # X and Y (step_length, cholesky_solve) and the diagonals update_cond_numbers reads
        ctx.schur_step(pX, pY, pXc, pYc, None, None, None, None, None)
        ctx.cholesky_diagonals(None, None, Sdiag, Qdiag)

    e2e_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - e0) / a.steps

    # ---- solve_schur_complement_equation on the resident factors ----
    # (called twice per Newton iteration: predictor and corrector)
    from sdpb_b200.synthetic import solve_rhs
    rx, ry = solve_rhs(prec, ctx.shapes, N, seed=7 + rank)
    dxh = pool.slab([x.shape for x in rx])
    dyh = pool.empty(ry.shape)
    solve_dev, solve_k = [], {}
    pdx = ptr_array(dxh)
    solve_wall = []
    for it in range(2 + a.steps):
        for dst, src in zip(dxh + [dyh], rx + [ry]):  # fresh right-hand sides (untimed)
            dst[...] = src
        barrier()
        s0 = time.perf_counter()
        ctx.solve_schur_complement_equation(pdx, dyh)  # synchronous: returns with dx, dy on the host
        if it >= 2:
            solve_wall.append(time.perf_counter() - s0)
            solve_dev.append(ctx.last_solve_ms())
            for name, ms in ctx.kernel_timings():
                solve_k.setdefault(name, []).append(ms)
    barrier()
    solve_api_s = float(np.mean(solve_wall))
    solve_dev_ms = float(np.mean(solve_dev))

    # ---- the round-1 contract for comparison: EVERY output copied back (L_j, L_j^-1 B_j, chol(Q)) ----
    e2e_all_s, d2h_all = None, None
    all_bytes = sum((s.schur_size ** 2 + N * s.schur_size) * ctx.ew * 8 for s in ctx.shapes)
    if not a.no_all_outputs and all_bytes < (8 << 30):  # pinned staging for L_j and L_j^-1 B_j
        Lh = pool.slab([(s.schur_size, s.schur_size, ctx.ew) for s in ctx.shapes])
        Ph = pool.slab([(N, s.schur_size, ctx.ew) for s in ctx.shapes])
        Qh = pool.empty((N, N, ctx.ew))
        pL, pP = ptr_array(Lh), ptr_array(Ph)
        d2h_all = sum(x.nbytes for x in Xc + Yc + Lh + Ph) + Qh.nbytes
        ctx.schur_step(pX, pY, pXc, pYc, None, None, pL, pP, Qh)
        barrier()
        e0 = time.perf_counter()
        for _ in range(a.steps):
            ctx.schur_step(pX, pY, pXc, pYc, None, None, pL, pP, Qh)
        barrier()
        e2e_all_s = (time.perf_counter() - e0) / a.steps

    # ---- compute_search_direction on the resident objects ----
    # (per Newton iteration: -XY + traces, R error, residues up, predictor, Frobenius products,
    # corrector, direction down -- what step.cxx:131-176 does around the two Schur solves)
    from sdpb_b200.synthetic import random_matrix as synth_matrix
    rng = np.random.Generator(np.random.PCG64(0x5D9B2000 + rank))
    pr = pool.slab([x.shape for x in sdp.X])
    for dst in pr:
        dst[...] = synth_matrix(rng, prec, dst.shape[1], dst.shape[0])
    dr = pool.slab([x.shape for x in rx])
    for dst in dr:
        dst[...] = synth_matrix(rng, prec, dst.shape[1], 1)
    prp = synth_matrix(rng, prec, N, 1)
    p_pr, p_dr = ptr_array(pr), ptr_array(dr)
    bm = np.zeros(ctx.ew, dtype=np.uint64)     # the packed element 0.125 (beta mu / mu stand-in)
    bm[0] = 1 << 32                            # sign +1, exponent 0
    bm[(prec + 63) // 64 + 2] = 1 << 61        # top limb
    dir_dev, dir_wall, dir_k = {}, [], {}
    sl_dev, sl_wall, sl_k, sl_iters = {}, [], {}, []
    dir_error = None
    for it in range(1 + a.steps):
        barrier()
        s0 = time.perf_counter()
        try:
            ctx.direction_begin()
        except Exception as e:  # the resident direction needs 7 more objects of the size of X: the largest
            dir_error = str(e)  # C5 corner runs the step but has no room for them
            break
        t_begin = ctx.last_direction_ms()
        k_begin = ctx.kernel_timings()
        ctx.direction_R_errors(bm)
        ctx.direction_set_residues(p_pr, p_dr, prp)
        ctx.compute_search_direction(bm, 0)
        t_pred = ctx.last_direction_ms()
        k_pred = ctx.kernel_timings()
        ctx.direction_frobenius()
        ctx.compute_search_direction(bm, 1)
        t_corr = ctx.last_direction_ms()
        wall_dir = time.perf_counter() - s0
        # row N3: step_length.cxx:27-46 for (X, dX) and (Y, dY) on the same resident objects
        ctx.step_length(0)
        t_slx = ctx.last_step_length_ms()
        k_sl = ctx.kernel_timings()
        ctx.step_length(1)
        t_sly = ctx.last_step_length_ms()
        wall_sl = time.perf_counter() - s0 - wall_dir
        if it >= 1:
            dir_wall.append(wall_dir)
            sl_wall.append(wall_sl)
            for k, v in (("minus_XY", t_begin), ("predictor", t_pred), ("corrector", t_corr)):
                dir_dev.setdefault(k, []).append(v)
            for k, v in (("primal", t_slx), ("dual", t_sly)):
                sl_dev.setdefault(k, []).append(v)
            for name, ms in k_begin + k_pred:
                dir_k.setdefault(name, []).append(ms)
            for name, ms in k_sl:
                sl_k.setdefault(name, []).append(ms)
            sl_iters = ctx.step_length_iterations()
    barrier()
    dir_api_s = float(np.mean(dir_wall)) if dir_wall else float("nan")
    sl_api_s = float(np.mean(sl_wall)) if sl_wall else float("nan")

    # ---- scale_multiply_add (-X Y and the other block GEMMs of step()) ----
    Ch = pool.slab([x.shape for x in sdp.X])
    sma_k = {}
    pC = ptr_array(Ch)
    sma_error = None
    try:
        ctx.scale_multiply_add(-1, pX, pY, 0, pC)
    except Exception as e:
        sma_error = str(e)
    barrier()
    s0 = time.perf_counter()
    for _ in range(a.steps if sma_error is None else 0):
        ctx.scale_multiply_add(-1, pX, pY, 0, pC)
        for name, ms in ctx.kernel_timings():
            sma_k.setdefault(name, []).append(ms)
    barrier()
    sma_api_s = (time.perf_counter() - s0) / a.steps if sma_error is None else float("nan")

    # ---- max over ranks -------------------------------------------------
    if world > 1:
        t = torch.tensor([ms_step, e2e_s, wall, solve_dev_ms, solve_api_s, e2e_all_s or 0.0], device="cuda",
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_s, wall, solve_dev_ms, solve_api_s, e2e_all_max = [float(x) for x in t.tolist()]
        if e2e_all_s is not None:
            e2e_all_s = e2e_all_max
    if rank != 0:
        pool.close()
        ctx.close()
        dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage ---------------------------------
    # A stage of SURVEY 8d's table is one label of the timeline; a label with sub-launches
    # ("trsm_Linv_B/gemm", "/diag": the levels of one batched triangular solve) is summed, and the
    # algorithmic bytes are the table's figure for the WHOLE stage -- (P_j^2/2 + 2 P_j N) E for the
    # solve, whatever the schedule re-reads -- divided by the launches it took.
    per_kernel = {k: (float(np.mean(v)) * len(v) / a.steps, len(v) // a.steps) for k, v in ktimes.items()}
    per_stage = {}
    for k, (ms, n) in per_kernel.items():
        g0 = per_stage.setdefault(k.split("/")[0], [0.0, 0])
        g0[0] += ms
        g0[1] += n
    dom = max(per_stage, key=lambda k: per_stage[k][0])
    dom_ms, dom_launches = per_stage[dom]
    peak, which = peaks()
    stage_bytes = algorithmic_bytes(dom, prec, ctx.shapes, N)
    abytes = stage_bytes / max(1, dom_launches)
    achieved = abytes / (dom_ms / max(1, dom_launches) * 1e-3) / 1e9
    traffic = recorded_traffic(a.workload, dom)
    roofline = {"bound": "hbm", "kernel": dom, "launches_per_step": dom_launches,
                "algorithmic_bytes_per_step": stage_bytes,
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic / max(1, dom_launches) if traffic else None,
                "peak_source": which, "share_of_step": dom_ms / serial_step, "ms_per_step": dom_ms,
                "timed_in": "single-stream pass (%.1f ms/step); the headline value overlaps independent "
                            "chains on side streams" % serial_step,
                "note": "bytes = SURVEY 8d's figure for the stage; multi-limb contraction: the INT32 multiply "
                        "pipe binds, not HBM (DESIGN.md §4), hence int_pipe beside it"}
    macs = limb_macs(dom, ctx.shapes, N)
    if macs:
        rate = macs * imad_per_mac(prec) / (dom_ms * 1e-3)
        roofline["int_pipe"] = {"unit": "IMAD.WIDE/s", "achieved": rate, "peak": IMAD_WIDE_PER_S,
                                "frac": rate / IMAD_WIDE_PER_S, "mpf_macs_per_launch": macs / max(1, dom_launches),
                                "imad_wide_per_mac": imad_per_mac(prec),
                                "peak_source": "measured, profiles/imad_rate4_r01.jsonl"}
    if a.kernels:
        for k, (ms, n) in sorted(per_stage.items(), key=lambda kv: -kv[1][0]):
            ab = algorithmic_bytes(k, prec, ctx.shapes, N)
            mc = limb_macs(k, ctx.shapes, N)
            log(f"  [{k:22s}] {ms:10.3f} ms/step  x{n}  {ab / 1e6:10.1f} MB  {ab / (ms * 1e-3) / 1e9 if ms else 0:8.1f} GB/s"
                f"  int_pipe {mc * imad_per_mac(prec) / (ms * 1e-3) / IMAD_WIDE_PER_S if ms else 0:6.3f}")
        for k, (ms, n) in sorted(per_kernel.items(), key=lambda kv: -kv[1][0]):
            if "/" in k:
                log(f"      {k:24s} {ms:10.3f} ms/step  x{n}")
        log(f"  stages(ms) {[round(x, 3) for x in stages]}")
        for k, v in solve_k.items():
            log(f"  [solve] {k:24s} {float(np.mean(v)):10.3f} ms")

    cpu = None
    solve_cpu = None
    sma_cpu = None
    if not a.no_cpu:
        try:
            sma_cpu = cpu_sma_sample(a.workload)
        except Exception as e:
            sma_cpu = {"value": None, "sample": f"unavailable: {e}"}
        try:
            solve_cpu = cpu_solve_sample(a.workload)
        except Exception as e:
            solve_cpu = {"value": None, "sample": f"unavailable: {e}"}
        try:
            cpu = cpu_full(a.workload, 1, 0, world=1, budget_s=60.0)
        except Exception as e:  # the product arm does not depend on the oracle
            cpu = {"value": None, "unit": "s/step", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}

    line = {"metric": "sec_per_newton_step_hot_path", "value": ms_step / 1e3, "unit": "s/step", "n_gpus": world,
            "steps": a.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "mpf%d" % prec, "data": "synthetic", "config": cfg,
            "clocks": clk.summary(), "wall_ms_per_step": wall / a.steps * 1e3,
            "e2e": {"value": e2e_s, "unit": "s/step", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "serial_ms_per_step": serial_step,
            "schur_solve": {"what": "solve_schur_complement_equation on the resident L_j, L_j^-1 B_j, chol(Q) "
                                    "(SURVEY 8f N1); called twice per Newton iteration",
                            "device_ms": solve_dev_ms, "api_ms_host_buffers": solve_api_s * 1e3,
                            "bytes_each_way": int(sum(x.nbytes for x in rx) + ry.nbytes),
                            "kernels_ms": {k: round(float(np.mean(v)), 4) for k, v in solve_k.items()},
                            "cpu": solve_cpu},
            "scale_multiply_add": {"what": "C = -X Y on the 2J PSD-shaped blocks through the C-ABI with host buffers "
                                           "(SURVEY 8f N2; scale_multiply_add.cxx:4-16)",
                                   "unavailable": sma_error,
                                   "api_ms_host_buffers": None if sma_error else sma_api_s * 1e3,
                                   "kernels_ms": {k: round(float(np.mean(v)), 4) for k, v in sma_k.items()},
                                   "bytes_h2d": int(2 * sum(x.nbytes for x in Xh)),
                                   "bytes_d2h": int(sum(x.nbytes for x in Ch)), "cpu": sma_cpu},
            "search_direction": {"what": "compute_search_direction.cxx:44-90 twice (predictor, corrector) plus "
                                         "-XY, traces, R error and Frobenius products, on device-resident X, Y, "
                                         "dX, dY, R, Z (SURVEY 8f rows N2); includes the two Schur solves; host "
                                         "traffic: residues up, per-block scalars down",
                                 "unavailable": dir_error,
                                 "api_ms_host_buffers": None if dir_error else dir_api_s * 1e3,
                                 "device_ms": {k: round(float(np.mean(v)), 3) for k, v in dir_dev.items()},
                                 "kernels_ms_predictor": {k: round(float(np.sum(v)) / a.steps, 4) for k, v in dir_k.items()},
                                 "bytes_h2d": int(sum(x.nbytes for x in pr) + sum(x.nbytes for x in dr) + prp.nbytes)},
            "step_length": {"what": "step_length.cxx:27-46 for (X, dX) and (Y, dY): congruence with the resident "
                                    "Cholesky factors, Householder tridiagonalisation, Laguerre min eigenvalue "
                                    "(SURVEY 8f row N3); 2J eigenvalues come down per call",
                            "api_ms_both_calls": None if dir_error else sl_api_s * 1e3,
                            "device_ms": {k: round(float(np.mean(v)), 3) for k, v in sl_dev.items()},
                            "kernels_ms_primal": {k: round(float(np.sum(v)) / a.steps, 4) for k, v in sl_k.items()},
                            "laguerre_steps_max": int(max(sl_iters)) if sl_iters else 0,
                            "laguerre_steps_mean": round(float(np.mean(sl_iters)), 2) if sl_iters else 0},
            "e2e_with_two_solves": {"what": "one Newton iteration's device work through the C-ABI with host "
                                            "buffers: the e2e step plus the predictor and corrector Schur solves",
                                    "value": e2e_s + 2 * solve_api_s, "unit": "s/iteration"},
            "e2e_newton_iteration": {"what": "everything of one Newton iteration that runs on the device, through the "
                                             "C-ABI with host buffers: the e2e step, the search direction (predictor "
                                             "and corrector, both Schur solves inside) and both step lengths "
                                             "(SURVEY 8 rows a + N1 + N2 + N3)",
                                     "value": None if dir_error else e2e_s + dir_api_s + sl_api_s, "unit": "s/iteration"},
            "e2e_all_outputs": {"what": "round-1 contract: sdpb_b200_schur_step with EVERY output copied back "
                                        "(X/Y factors, L_j, L_j^-1 B_j, chol(Q))",
                                "value": e2e_all_s, "unit": "s/step", "d2h_bytes_per_step": d2h_all},
            "numa": numa,
            "stages_ms": {n: round(float(v), 4) for n, v in zip(
                ["chol_XY", "pairings", "schur_assembly", "chol_S+trsm", "normalize", "exact_syrk", "restore",
                 "chol_Q", "step"], stages)}}
    print(json.dumps(line), flush=True)
    pool.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _mix(shapes):
    c = {}
    for s in shapes:
        c[s] = c.get(s, 0) + 1
    return ", ".join(f"{n}x(m={m},n={nn})" for (m, nn), n in c.items())


if __name__ == "__main__":
    main()
