"""mpfx (the host/device fixed-limb arithmetic) against the real libgmp.

Compiles tests/cpp/mpfx_fuzz.cpp for the host and runs it: every mpf operation
the hot path uses (mul, add, sub, div, sqrt, 2^k scalings, /4, cmp, integer
truncation) on variable-size and adversarial operands, at eight precisions.
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    exe = os.path.join(ROOT, "build", "mpfx_fuzz")
    src = os.path.join(ROOT, "tests", "cpp", "mpfx_fuzz.cpp")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, src, "-l:libgmp.so.10"], check=True)
    return exe


def test_mpfx_matches_libgmp_bit_for_bit():
    exe = _build()
    for seed in (1, 2, 3):
        r = subprocess.run([exe, "6000", str(seed)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "0 mismatches" in r.stdout


def test_mpfw_register_form_matches_mpfx():
    """mpfw (the register-resident hot-loop arithmetic: short-product multiply,
    unified add/sub, division by reciprocal, Newton sqrt / reciprocal) against
    mpfx on random and adversarial operands; the rare exact-fallback paths must
    be exercised and the Newton paths must never need their slow fallback."""
    exe = os.path.join(ROOT, "build", "mpfw_fuzz")
    src = os.path.join(ROOT, "tests", "cpp", "mpfw_fuzz.cpp")
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    r = subprocess.run([exe, "12000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MISMATCH" not in r.stdout
