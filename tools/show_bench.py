#!/usr/bin/env python3
"""One line per workload of a bench.py JSON line: tools/show_bench.py <file>"""
import json
import sys

l = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])


def show(name, r):
    if "error" in r:
        print(name, "ERROR", r["error"])
        return
    e = r["e2e"]
    print(f"{name}: kernel {r['value']:.1f} GCUPS ({r['ms_per_step']:.2f} ms)  e2e {e.get('value') or 0:.1f} ({e.get('ms_per_step', 0):.1f} ms, h2d {e.get('h2d_bytes_per_step', 0) / 1e6:.0f} MB)"
          f"  ascii {e.get('ascii_input', {}).get('value', 0):.1f}  frac {r['roofline']['frac']:.3f}  cpu {r.get('cpu_baseline', {}).get('value', 0):.1f}"
          f"  parity {r.get('parity', {}).get('mismatches')}/{r.get('parity', {}).get('pairs_checked')}")


show(l["config"]["workload"].split(":")[0], l)
if "strong_scaling" in l:
    print("  strong:", {k: v for k, v in l["strong_scaling"].items() if k in ("n_gpus", "value", "ms_per_step", "parity", "error")})
if "pageable_caller_buffers" in l["e2e"]:
    print("  pageable:", l["e2e"]["pageable_caller_buffers"]["value"])
for k, v in l.get("configs", {}).items():
    show(k, v)
