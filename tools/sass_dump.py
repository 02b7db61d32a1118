#!/usr/bin/env python3
"""Print the SASS of one address range of one kernel, with an opcode histogram.
usage: tools/sass_dump.py <lib.so> <kernel name substring> <from hex> <to hex> [-q]"""
import collections
import re
import subprocess
import sys

lib, pat, a, b = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on, hist = False, collections.Counter()
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        on = pat in m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if on and m and a <= int(m.group(1), 16) <= b:
        t = m.group(2).strip()
        if "-q" not in sys.argv:
            print(m.group(1), t)
        hist[re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]] += 1
print(sum(hist.values()), dict(hist.most_common()))
