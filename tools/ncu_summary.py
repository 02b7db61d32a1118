#!/usr/bin/env python3
"""Append one capture to profiles/r02_ncu_kernel_summary.json (BA_NCU_SUMMARY overrides the file name) from `ncu -i X.ncu-rep --page raw --csv` output.

usage: tools/ncu_summary.py raw.csv "<capture name>" <pairs in launch> "<note>"
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))


def f(k, default=None):
    try:
        return float(m[k])
    except Exception:
        return default


pairs = int(sys.argv[3])
stall = {}
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
        stall[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = f(k)
cap = {
    "capture": sys.argv[2], "pairs_in_launch": pairs, "note": sys.argv[4] if len(sys.argv) > 4 else "",
    "kernel": m.get("Kernel Name", ""),
    "gpu__time_duration_ms": f("gpu__time_duration.sum") / (1e6 if f("gpu__time_duration.sum") > 1e5 else 1.0),
    "registers_per_thread": f("launch__registers_per_thread"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "pipe_lsu_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "icc_hit_rate_pct": f("sm__icc_request_hit_rate.pct"),
    "inst_executed": f("smsp__inst_executed.sum"),
    "dram_bytes_read": f("dram__bytes_read.sum"), "dram_bytes_write": f("dram__bytes_write.sum"),
    "dram_bytes_unit": "as printed by ncu (see dram_bytes_per_pair for bytes)",
    "stall_per_issue": {k: v for k, v in sorted(stall.items(), key=lambda kv: -(kv[1] or 0))[:9]},
}
# ncu prints dram bytes in the unit of the second header row
units = dict(zip(hdr, rows[1]))
def to_bytes(k):
    return f(k) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units.get(k, "byte"), 1.0)


cap["dram_bytes_read"], cap["dram_bytes_write"] = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
cap["dram_bytes_unit"] = "bytes"
cap["dram_bytes_per_pair"] = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) / pairs
cap["warp_instructions_per_pair"] = cap["inst_executed"] / pairs
tu = units.get("gpu__time_duration.sum", "")
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, None)
if tscale is not None:
    cap["gpu__time_duration_ms"] = f("gpu__time_duration.sum") * tscale
path = os.path.join(ROOT, "profiles", os.environ.get("BA_NCU_SUMMARY", "r02_ncu_kernel_summary.json"))
doc = json.load(open(path)) if os.path.exists(path) else {"note": "one entry per ncu --set full capture (tools/ncu_job.sh); numbers under the profiler are not bench values", "captures": []}
doc["captures"] = [c for c in doc["captures"] if c["capture"] != cap["capture"]] + [cap]
json.dump(doc, open(path, "w"), indent=1)
print(json.dumps(cap, indent=1))
