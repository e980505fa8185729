"""Markdown tables of DESIGN.md §9 from the round's bench lines under profiles/.
usage: python tools/results_table.py v5   (reads profiles/bench_r02_<tag>*.json, profiles/scale_r02_*.json)"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "v5"


def load(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def row(name, d):
    r = d.get("roofline") or {}
    ip = r.get("int_pipe") or {}
    st = d.get("stages_ms") or {}
    top = sorted(((v, k) for k, v in st.items() if k != "step"), reverse=True)[:3]
    return "| %s | %d b | %.1f | %.1f | %s %.4f | %s | %s |" % (
        name, d["config"]["precision_bits"], d["ms_per_step"], d["e2e"]["value"] * 1e3, r.get("kernel", "-"),
        r.get("frac", 0), ("%.2f" % ip["frac"]) if ip.get("frac") else "-",
        ", ".join("%s %.0f" % (k, v) for v, k in top))


print("| workload | prec | ms/step (device) | e2e ms (C-ABI, host buffers) | dominant stage, HBM frac | int_pipe | largest stages (ms) |")
print("|---|---|---|---|---|---|---|")
main = load(os.path.join(ROOT, "profiles", "bench_r02_%s.json" % tag))
if main:
    print(row("c3 (J=600, N=300)", main))
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r02_%s_*.json" % tag))):
    name = os.path.basename(path)[len("bench_r02_%s_" % tag):-5]
    if name == "ref":
        continue
    d = load(path)
    if d and "ms_per_step" in d:
        print(row(name, d))
print()
print("| GPUs | workload | ms/step (device, max over ranks) | e2e ms | efficiency (device / e2e) |")
print("|---|---|---|---|---|")
base = {}
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "scale_r02_%s_*.json" % tag)),
                   key=lambda p: (p.split("_")[-2], int(p.split("_n")[-1][:-5]))):
    d = load(path)
    if not d or "ms_per_step" not in d:
        continue
    w = d["config"]["workload"].split(":")[0]
    n = d["n_gpus"]
    if n == 1 or w not in base:
        base.setdefault(w, (d["ms_per_step"], d["e2e"]["value"] * 1e3))
    b = base[w]
    print("| %d | %s | %.1f | %.1f | %.3f / %.3f |" % (n, w, d["ms_per_step"], d["e2e"]["value"] * 1e3,
                                                      b[0] / d["ms_per_step"], b[1] / (d["e2e"]["value"] * 1e3)))
if main:
    c = main.get("cpu_baseline") or {}
    print()
    print("CPU arm (c3, full block list, measured): %.2f s/step on %s cores (%s); e2e ratio %.0fx" % (
        c.get("value") or float("nan"), c.get("cores"), c.get("kind"), (c.get("value") or 0) / main["e2e"]["value"]))
    for k in ("schur_solve", "search_direction", "step_length", "scale_multiply_add"):
        v = main.get(k) or {}
        print(k, {kk: vv for kk, vv in v.items() if kk in ("device_ms", "api_ms_host_buffers", "api_ms_both_calls",
                                                            "laguerre_steps_mean", "laguerre_steps_max")})
