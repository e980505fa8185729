"""Compare a solver output directory with a reference golden `out/` directory the
way the reference's own end-to-end test does (reference
test/src/integration_tests/util/diff_sdpb_out.cxx:194-285, test/src/test_util/diff.hxx:50-75):
terminateReason equal, primalObjective / dualObjective, every x_j / y / z entry, c_minus_By and
every field of every iteration of iterations.json (times and block_name excluded, errors below
2^-(diff_precision/2) skipped) within  |a-b| < 2^-99 (|a|+|b|)."""
import json
import os
import re

import mpmath

DIFF_PRECISION = 99  # end-to-end.test.cxx:26-27


def _mp():
    mpmath.mp.prec = 1100


def close(a, b, bits=DIFF_PRECISION):
    _mp()
    a, b = mpmath.mpf(a), mpmath.mpf(b)
    if a == b:
        return True
    return abs(a - b) < mpmath.ldexp(1, -bits) * (abs(a) + abs(b))


def parse_out_txt(path):
    d = {}
    for line in open(path):
        m = re.match(r"\s*([^=]+?)\s*=\s*(.*?);\s*$", line)
        if m:
            d[m.group(1).strip()] = m.group(2).strip().strip('"')
    return d


def parse_vector(path):
    lines = open(path).read().split()
    h, w = int(lines[0]), int(lines[1])
    vals = lines[2:]
    assert len(vals) == h * w, path
    return (h, w), vals


def diff_out_dirs(ours, golden, keys=("terminateReason", "primalObjective", "dualObjective"),
                  files=None, iterations_name="iterations.json", max_iterations=None):
    """Returns a list of mismatch descriptions (empty = pass)."""
    bad = []
    a, b = parse_out_txt(os.path.join(ours, "out.txt")), parse_out_txt(os.path.join(golden, "out.txt"))
    for k in keys:
        if k == "terminateReason":
            if a[k] != b[k]:
                bad.append(f"out.txt {k}: {a[k]!r} != {b[k]!r}")
        elif not close(a[k], b[k]):
            bad.append(f"out.txt {k}: {a[k][:40]} vs {b[k][:40]}")
    if files is None:
        files = sorted(f for f in os.listdir(golden) if re.match(r"(x_\d+|y|z)\.txt$", f))
    for f in files:
        if not os.path.exists(os.path.join(ours, f)):
            bad.append(f"missing {f}")
            continue
        (sa, va), (sb, vb) = parse_vector(os.path.join(ours, f)), parse_vector(os.path.join(golden, f))
        if sa != sb:
            bad.append(f"{f}: shape {sa} != {sb}")
            continue
        for i, (p, q) in enumerate(zip(va, vb)):
            if not close(p, q):
                bad.append(f"{f}[{i}]: {p[:40]} vs {q[:40]}")
                break
    cg = os.path.join(golden, "c_minus_By", "c_minus_By.json")
    co = os.path.join(ours, "c_minus_By", "c_minus_By.json")
    if os.path.exists(cg) and max_iterations is None:
        ja, jb = json.load(open(co))["c_minus_By"], json.load(open(cg))["c_minus_By"]
        if [len(x) for x in ja] != [len(x) for x in jb]:
            bad.append("c_minus_By: shapes differ")
        else:
            for j, (ba, bb) in enumerate(zip(ja, jb)):
                for i, (p, q) in enumerate(zip(ba, bb)):
                    if not close(p, q):
                        bad.append(f"c_minus_By[{j}][{i}]: {p[:40]} vs {q[:40]}")
                        break
    if iterations_name is not None:  # None: a run restarted from a checkpoint has its own iteration history
        bad += diff_iterations(os.path.join(ours, "iterations.json"), os.path.join(golden, iterations_name),
                               max_iterations)
    return bad


def diff_iterations(ours, golden, max_iterations=None):
    bad = []
    ia, ib = json.load(open(ours)), json.load(open(golden))
    if max_iterations is not None:
        ia, ib = ia[:max_iterations], ib[:max_iterations]
    if len(ia) != len(ib):
        bad.append(f"iterations.json: {len(ia)} iterations vs {len(ib)}")
    _mp()
    abs_eps = mpmath.ldexp(1, -(DIFF_PRECISION // 2))
    for ra, rb in zip(ia, ib):
        if list(ra.keys()) != list(rb.keys()):
            bad.append(f"iteration {rb['iteration']}: keys differ")
            continue
        for k in ra:
            if k in ("total_time", "iter_time", "block_name"):
                continue
            if k == "iteration":
                if ra[k] != rb[k]:
                    bad.append(f"iteration number {ra[k]} vs {rb[k]}")
                continue
            if k in ("P-err", "p-err", "D-err", "R-err"):
                if abs(mpmath.mpf(ra[k])) + abs(mpmath.mpf(rb[k])) < abs_eps:
                    continue
            if not close(ra[k], rb[k]):
                bad.append(f"iteration {rb['iteration']} {k}: {ra[k][:40]} vs {rb[k][:40]}")
    return bad
