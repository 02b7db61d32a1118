#!/usr/bin/env python3
"""Emulator soak aimed at the packed rectangles taller than 256 rows (place_rect_pk_tall): min block 32 / 64 (the kernels
that carry that code), max block 1024..8192, big indels that force the block to grow, all (TRACE, X_DROP) combinations,
nucleotide / amino-acid / byte scoring. Compared with the oracle like tools/soak.py; also reports how many cells went
through packed and through exact rectangles.   usage: tools/soak_tall.py <worker id> <minutes>
Test infrastructure (uses tests/ and oracle/)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import backend, parity
from block_aligner_b200 import api, workloads
P = workloads.params
wid = int(sys.argv[1]); minutes = float(sys.argv[2])
rng = np.random.default_rng(500 + wid)
lib = backend.emu_lib(); al = api.Aligner(lib)
lib.L.ba_emu_stats.argtypes = [C.c_void_p, C.c_int]
def stats():
    out = (C.c_uint64 * 4)(); lib.L.ba_emu_stats(out, 1)
    return out[0], out[1]
t_end = time.time() + 60 * minutes
n = bad_total = 0; pk = ex = 0
while time.time() < t_end:
    lo = int(rng.choice([32, 64])); hi = int(rng.choice([1024, 2048, 2048, 4096, 8192]))
    flags = int(rng.choice([0, api.XDROP, api.TRACE, api.TRACE | api.XDROP, api.TRACE | api.XDROP]))
    kind = int(rng.choice([0, 0, 1, 2]))
    if kind == 0:
        sc, mat, gaps, alpha = api.SCORING_NUC, ("NW1" if rng.random() < .5 else (2, -4)), None, 0
        gaps = (-2, -1) if mat == "NW1" else (-6, -2)
    elif kind == 1:
        sc, mat, gaps, alpha = api.SCORING_AA, "BLOSUM62", (-11, -1), 1
    else:
        sc, mat, gaps, alpha = api.SCORING_BYTE, (1, -1), (-2, -1), 0
        if (flags & api.TRACE) and (flags & api.XDROP):
            flags &= ~api.XDROP      # no reference CIGAR exists for ByteMatrix + X_DROP ends in the padding (see tools/soak.py)
    rate = float(rng.choice([0.02, 0.05, 0.12]))
    lmin = int(rng.integers(1500, 5000))
    w = dict(scoring=sc, matrix=mat, gaps=gaps, size=(lo, hi), x_drop=int(rng.choice([50, 200, 400])), flags=flags, stream=int(rng.integers(1, 90)),
             gen=P(alphabet=alpha, len_dist=0, len_min=lmin, len_max=lmin + int(rng.integers(0, 4000)), sub_rate=rate, ins_rate=rate / 2, del_rate=rate / 2,
                   long_indel_mean=float(rng.choice([0.0, 1.0, 3.0])), long_indel_len=float(rng.choice([60.0, 200.0])), suffix_len=int(rng.choice([0, 300])),
                   big_indel_prob=float(rng.choice([0.5, 0.9])), big_indel_min=int(rng.choice([100, 400, 1000])), big_indel_max=int(rng.choice([1500, 3000]))))
    seed = int(rng.integers(1 << 30))
    stats()
    bad = parity.check_workload(lib, al, w, 8, seed=seed)
    a, b = stats(); pk += a; ex += b
    if bad: print("MISMATCH", w, seed, flush=True)
    bad_total += bad; n += 8
print(f"tall soak worker {wid}: {n} pairs, {bad_total} mismatching, packed-rectangle cells {pk}, exact cells {ex}", flush=True)
