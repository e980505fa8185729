"""Pins the oracle (and, on the GPU box, the CUDA path) against the reference's own
golden fixtures: the full Newton trajectories of its end-to-end tests
(reference test/src/integration_tests/cases/end-to-end.test.cxx:180-381), packed
byte for byte into tests/golden/*.tar.gz by tests/golden/make_fixtures.py.

The host solver (sdpb_b200/csrc/host/solver.hpp, SDP_Solver::run/step) is run with
the reference test's sdpb options on the fixture's sdp/ directory; out.txt, x_j, y, z,
c_minus_By and EVERY field of EVERY iteration of iterations.json (mu, objectives, gap,
errors, step lengths, beta, Q_cond_number, max_block_cond_number) must agree with the
reference's files within the reference's own tolerance  |a-b| < 2^-99 (|a|+|b|)
(test/src/test_util/diff.hxx:50-75; tests/golden_check.py restates diff_sdpb_out.cxx).

  not gpu : hot path = CPU oracle (oracle/hotpath_core.hpp on libgmp)  -> the oracle is pinned
  gpu     : hot path = sm_100a kernels through the C-ABI (libsdpb_b200_host.so -> libsdpb_b200.so)
"""
import ctypes
import json
import os
import tarfile

import pytest

import golden_check

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = json.load(open(os.path.join(GOLDEN, "cases.json")))


def _unpack(name, tmp_path):
    with tarfile.open(os.path.join(GOLDEN, name + ".tar.gz")) as tar:
        tar.extractall(tmp_path, filter="data")
    return os.path.join(tmp_path, name)


def _solve(lib, fn, args):
    argv = (ctypes.c_char_p * len(args))(*[a.encode() for a in args])
    buf = ctypes.create_string_buffer(8192)
    f = getattr(lib, fn)
    f.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_char_p, ctypes.c_size_t]
    f.restype = ctypes.c_int
    rc = f(len(args), argv, buf, 8192)
    assert rc == 0, buf.value.decode()
    return json.loads(buf.value.decode())


def _run_case(lib, fn, name, tmp_path):
    case = CASES[name]
    root = _unpack(name, str(tmp_path))
    out = os.path.join(str(tmp_path), "ours")
    os.makedirs(out, exist_ok=True)
    args = ["--sdpDir", os.path.join(root, "sdp"), "--outDir", out, "--precision", str(case["precision"])]
    summary = _solve(lib, fn, args + case["sdpb_args"])
    kw = {"iterations_name": case["iterations"]}
    if case["out_txt_keys"]:
        kw["keys"] = tuple(case["out_txt_keys"])
    bad = golden_check.diff_out_dirs(out, os.path.join(root, "out"), **kw)
    assert not bad, f"{name}: {len(bad)} mismatches vs the reference golden, first: {bad[:5]}"
    n_golden = len(json.load(open(os.path.join(root, "out", case["iterations"]))))
    assert summary["iterations"] == n_golden
    return summary


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_replays_reference_golden_trajectory(name, tmp_path, oracle):
    lib = oracle.load_oracle()
    s = _run_case(lib, "oracle_solve", name, tmp_path)
    assert s["hot_path"].startswith("cpu-oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_b200_replays_reference_golden_trajectory(name, tmp_path):
    lib = ctypes.CDLL(os.path.join(ROOT, "sdpb_b200", "libsdpb_b200_host.so"))
    s = _run_case(lib, "sdpb_b200_solve", name, tmp_path)
    assert s["hot_path"].startswith("sm_100a")


def test_host_solver_library_exports_declared_symbols_and_has_no_cpu_hot_path():
    """include/sdpb_b200_solver.h <-> libsdpb_b200_host.so; the product library must not
    contain (or link) the oracle."""
    import re
    import subprocess
    text = open(os.path.join(ROOT, "include", "sdpb_b200_solver.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(sdpb_b200_[a-z_0-9]+)\s*\(", text)))
    assert names == ["sdpb_b200_sdp_to_binary", "sdpb_b200_solve"]
    path = os.path.join(ROOT, "sdpb_b200", "libsdpb_b200_host.so")
    lib = ctypes.CDLL(path)
    for n in names:
        assert hasattr(lib, n)
    syms = subprocess.run(["nm", "-D", path], capture_output=True, text=True).stdout
    assert "oracle_" not in syms
    deps = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "liboracle" not in deps and "libsdpb_b200.so" in deps
