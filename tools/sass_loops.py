#!/usr/bin/env python3
"""Static view of a kernel's loops from `cuobjdump -sass`: for every backward branch, the size of the loop body and how
many DPX / shuffle / shared / local-memory (spill) instructions it holds. Used to check before spending GPU time that a
change keeps the hot loops free of spills and small enough for the instruction cache.
usage: tools/sass_loops.py <lib.so> <kernel name substring> [min DPX per loop]"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
min_dpx = int(sys.argv[3]) if len(sys.argv) > 3 else 1
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, kernels = None, {}
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        kernels[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in kernels.items():
    if pat not in name:
        continue
    total = len(ins)
    nl = sum(1 for _, t in ins if re.search(r"\b(LDL|STL)", t))
    print(f"== {name}: {total} instructions ({total * 16 / 1024:.1f} KB), {nl} local-memory instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr:
                loops.append((addr[tgt], i))
    for lo, hi in sorted(loops):
        body = [t for _, t in ins[lo:hi + 1]]
        dpx = sum(1 for t in body if re.search(r"VIADDMNMX|VIMNMX", t))
        if dpx < min_dpx:
            continue
        c = lambda rx: sum(1 for t in body if re.search(rx, t))
        print(f"  loop @{ins[lo][0]:#07x}..{ins[hi][0]:#07x}: {hi - lo + 1:5d} instr, DPX {dpx:4d} (x2 {c(r'S16x2|16x2')}), SHFL {c('SHFL')}, "
              f"LDS/STS {c(r'LDS|STS')}, LDG/STG {c(r'LDG|STG|LD.E|ST.E')}, local {c(r'LDL|STL')}, PRMT {c('PRMT')}, CALL {c('CALL')}")
