"""profiles/traffic_r02_<tag>_trsm.json (+ _syrk.json) from the ncu CSV of tools/gpu/session_r2_<tag>.sh:
dram__bytes_read.sum + dram__bytes_write.sum of the launches of ONE c3 step behind the timeline labels
trsm_Linv_B and syrk_imma_kernel, which bench.py's roofline.traffic reads.
usage: python tools/traffic_record.py v5"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "v5"
src = os.path.join(ROOT, "profiles", "traffic_r02_%s.csv" % tag)
lines = [l for l in open(src) if l.startswith('"')]
by = collections.OrderedDict()
for x in csv.DictReader(lines):
    rec = by.setdefault(x["ID"], {"kernel": x["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0],
                                  "grid": x["Grid Size"]})
    rec[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
recs = list(by.values())


def emit(label, picked, note):
    launches = [{"kernel": b["kernel"], "grid": b["grid"], "ms": round(b["gpu__time_duration.sum"] / 1e6, 3),
                 "dram_read_MB": round(b["dram__bytes_read.sum"] / 1e6, 1),
                 "dram_write_MB": round(b["dram__bytes_write.sum"] / 1e6, 1)} for b in picked]
    tot = sum(b["dram__bytes_read.sum"] + b["dram__bytes_write.sum"] for b in picked)
    out = os.path.join(ROOT, "profiles", "traffic_r02_%s_%s.json" % (tag, label.split("_")[0]))
    json.dump({"workload": "c3", "label": label,
               "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control "
                         "none -k regex:trsm_|syrk_ (SDPB_B200_CONCURRENCY=0) python bench.py --steps 1 --warmup 3 --no-cpu "
                         "--no-all-outputs; tools/gpu/session_r2_%s.sh; %s" % (tag, note),
               "dram_bytes": tot, "launches": launches}, open(out, "w"), indent=1)
    print(label, len(picked), "launches", round(tot / 1e9, 3), "GB", round(sum(l["ms"] for l in launches), 3), "ms")


# one step's L_j^-1 B_j: the consecutive trsm launches whose grids count 600 or 150 blocks (one group:
# 7 update + 8 diagonal levels = 15), or 150 and then 450 blocks (split by size class: 15 + 5 = 20)
is_B = lambda b: b["kernel"].startswith("trsm_") and b["grid"].split(",")[0].strip("( ") in ("600", "150", "450")
want = 20 if any(is_B(b) and b["grid"].split(",")[0].strip("( ") == "450" for b in recs) else 15
start = next(i for i, b in enumerate(recs) if is_B(b))
step = []
for b in recs[start:]:
    if is_B(b):
        step.append(b)
    elif step and b["kernel"].startswith("trsm_"):
        break
    if len(step) == want:
        break
assert len(step) == want, len(step)
emit("trsm_Linv_B", step, "the %d launches of one step" % want)
syrk = [b for b in recs if b["kernel"] == "syrk_imma_kernel"][:1]
pack = [b for b in recs if b["kernel"] == "syrk_pack_kernel"][:1]
if syrk:
    emit("syrk_imma_kernel", syrk, "one launch")
if pack:
    emit("syrk_pack_kernel", pack, "one launch")
