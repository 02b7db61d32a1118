#!/usr/bin/env python3
"""Instruction / stall-sample split per device function from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.

usage: tools/ncu_source_split.py src.csv [top_lines]
Function extents are read from the current sources (block_aligner_b200/csrc), so run it on a capture of the same
revision. Inlined wrappers (ba_warp.cuh, CUDA headers) are reported per file.
"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "block_aligner_b200", "csrc")


def function_ranges(path):
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:BA_DEV|BA_HD|BA_DEV_NOINLINE|__global__|static)\b[^;{]*?\b(\w+)\s*\(")
    starts = []
    lines = open(path).read().split("\n")
    for i, l in enumerate(lines, 1):
        if l.startswith(" ") or l.startswith("//"):
            continue
        m = pat.match(l)
        if m:
            starts.append((i, m.group(1)))
    return [(a, (starts[k + 1][0] - 1 if k + 1 < len(starts) else len(lines)), name) for k, (a, name) in enumerate(starts)]


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, data = None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0].isdigit():
        try:
            data.append((cur, int(r[0]), int(r[7]), int(r[4]), r[1].strip()[:110]))
        except (ValueError, IndexError):
            pass
tot = sum(d[2] for d in data) or 1
ts = sum(d[3] for d in data) or 1
print(f"total warp instructions {tot}, stall samples {ts}")
agg = collections.Counter()
samp = collections.Counter()
ranges = {f: function_ranges(os.path.join(CSRC, f)) for f in ("ba_kernel.cuh", "ba_packed.cuh", "ba_runtime.cu") if os.path.exists(os.path.join(CSRC, f))}
for f, line, inst, s, _ in data:
    name = f
    for a, b, fn in ranges.get(f, []):
        if a <= line <= b:
            name = f"{f}:{fn}"
            break
    agg[name] += inst
    samp[name] += s
print("%-44s %8s %8s" % ("function (or file, for inlined wrappers)", "inst %", "samples %"))
for name, v in agg.most_common(24):
    print("%-44s %8.1f %8.1f" % (name, 100 * v / tot, 100 * samp[name] / ts))
print()
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%-14s %5d inst %5.2f%% samples %5.2f%%  %s" % (d[0], d[1], d[2] / tot * 100, d[3] / ts * 100, d[4]))
