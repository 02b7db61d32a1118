#!/usr/bin/env python3
"""Latency of the reference-compatible per-pair calls (Part 1 of the C ABI: block_align_aa*, one pair per call = one
batch of one on the GPU) next to the reference's own published per-call times (3.9 us at
length 100, 23 us at length 1000 for its protein scan, block 32..=32 / 32..=256 on AVX2). Tells drop-in users when to move to
the batch API. usage: tools/part1_latency.py [calls]   -> one JSON line"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from block_aligner_b200 import api  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
lib = api.Library()
b62 = C.addressof((C.c_int8 * 864).in_dll(lib.L, "BLOSUM62"))
rng = np.random.default_rng(1)
out = {}
for length, size in ((100, (32, 32)), (1000, (32, 256)), (1000, (32, 32))):
    r = bytes(rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8), length))
    q = bytearray(r)
    for p in rng.integers(0, length, size=length // 10):
        q[int(p)] = ord("W")
    pq, pr = api.PaddedBytes(lib, bytes(q), size[1]), api.PaddedBytes(lib, r, size[1])
    for trace in (False, True):
        blk = api.Block(lib, length, length, size[1], trace=trace)
        cg = api.Cigar(lib, length, length)
        for _ in range(20):
            blk.align(pq, pr, b62, (-11, -1), size, 0)
        t0 = time.perf_counter()
        for _ in range(calls):
            res = blk.align(pq, pr, b62, (-11, -1), size, 0)
            if trace:
                blk.cigar(res[1], res[2], cg)
        dt = (time.perf_counter() - t0) / calls
        out[f"len{length}_block{size[0]}-{size[1]}" + ("_trace+cigar" if trace else "")] = round(dt * 1e6, 1)
print(json.dumps({"part1_us_per_call": out, "calls": calls,
                  "reference_published_us": {"len100_block32-32": 3.9, "len1000": 23.0, "len100_trace (Block allocated per call)": 355.0},
                  "note": "block_align_aa / block_align_aa_trace + block_cigar_aa_trace through ctypes; every call is a batch of one: H2D, two kernel launches, D2H and a stream synchronisation"}))
