#!/usr/bin/env python3
"""Randomised soak of the device source (emulated on the CPU) against the oracle: many small batches over random
scoring kinds, flag combinations, block ranges, length ranges and error profiles. Every pair is compared on score,
end indices, CIGAR and the path-determined cell count (tests/parity.py).

usage: tools/emu_soak.py <minutes> <worker id> [pairs per batch]      -> one summary line on stdout
Test infrastructure (uses tests/emu and oracle/); the GPU runs the same source, see tests/test_gpu_parity.py.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import backend  # noqa: E402
import parity  # noqa: E402
from block_aligner_b200 import api, workloads  # noqa: E402

minutes, wid = float(sys.argv[1]), int(sys.argv[2])
per = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rng = np.random.default_rng(1000 + wid)
lib = backend.emu_lib()
al = api.Aligner(lib)
P = workloads.params
t_end = time.time() + 60 * minutes
n_pairs = n_bad = n_batches = 0
cells = 0
kinds = {}
while time.time() < t_end:
    scoring = int(rng.choice([api.SCORING_NUC, api.SCORING_NUC, api.SCORING_AA, api.SCORING_BYTE, api.SCORING_PROFILE]))
    flags = int(rng.choice([0, api.XDROP, api.TRACE, api.TRACE | api.XDROP]))
    ext = 0
    if scoring != api.SCORING_PROFILE and rng.random() < 0.2:
        ext = int(rng.choice([api.LOCAL_START, api.FREE_QUERY_START_GAPS]))
    lo = int(rng.choice([16, 32, 32, 64, 128]))
    hi = int(rng.choice([s for s in (16, 32, 64, 128, 256, 512, 1024, 2048) if s >= lo]))
    lmin = int(rng.integers(1, 400))
    lmax = lmin + int(rng.integers(0, 1500))
    rate = float(rng.choice([0.0, 0.02, 0.05, 0.12, 0.3]))
    gen = P(alphabet=0 if scoring in (api.SCORING_NUC, api.SCORING_BYTE) else 1, len_dist=0, len_min=lmin, len_max=lmax,
            sub_rate=rate, ins_rate=rate / 2, del_rate=rate / 2, long_indel_mean=float(rng.choice([0.0, 1.0, 3.0])),
            long_indel_len=float(rng.choice([10.0, 60.0, 200.0])), suffix_len=int(rng.choice([0, 50, 300])),
            big_indel_prob=float(rng.choice([0.0, 0.5])), big_indel_min=50, big_indel_max=400)
    if scoring == api.SCORING_NUC:
        matrix, gaps = (("NW1", (-2, -1)) if rng.random() < 0.5 else ((int(rng.integers(1, 6)), -int(rng.integers(1, 7))), None))
        if gaps is None:
            e = -int(rng.integers(1, 4)); gaps = (e - int(rng.integers(1, 9)), e)
    elif scoring == api.SCORING_AA:
        matrix, gaps = str(rng.choice(["BLOSUM62", "BLOSUM45", "PAM250"])), (-11, -1)
    elif scoring == api.SCORING_BYTE:
        matrix, gaps = (1, -1), (-2, -1)
    else:
        matrix, gaps = None, None
    w = dict(scoring=scoring, matrix=matrix, gaps=gaps, size=(lo, hi), x_drop=int(rng.choice([0, 10, 50, 400])), flags=flags | ext,
             stream=int(rng.integers(1, 1000)), gen=gen)
    if scoring == api.SCORING_BYTE:
        w["matrix"] = None
        # ByteMatrix pads with byte 0, which *matches* the other sequence's padding, so an X-drop alignment can end
        # past both sequences; the reference's traceback then panics ("Traceback cigar end position must be in
        # bounds!", scan_block.rs:1483) -- there is no CIGAR to compare. Keep TRACE, drop X_DROP for this kind.
        if (w["flags"] & api.TRACE) and (w["flags"] & api.XDROP):
            w["flags"] &= ~api.XDROP
            flags &= ~api.XDROP
    try:
        if scoring == api.SCORING_BYTE:
            qa, qo, ra, ro = workloads.generate(gen, per, seed=int(rng.integers(1 << 30)), stream=w["stream"])
            m = np.array([1, -1], dtype=np.int8)
            got = parity.run_lib(lib, al, scoring, m, gaps, w["size"], w["x_drop"], w["flags"], bool(flags & api.TRACE), qa, qo, ra, ro)
            exp = parity.oracle_batch(scoring, m, gaps, w["size"], w["x_drop"], w["flags"], bool(flags & api.TRACE), qa, qo, ra, ro)
            bad = parity.compare("soak-byte", got, exp)
            cells += int(exp[1].sum())
        else:
            bad = parity.check_workload(lib, al, w, per, seed=int(rng.integers(1 << 30)))
    except api.BlockAlignerError as e:      # argument combinations the reference rejects too (e.g. open >= extend)
        continue
    n_batches += 1
    n_pairs += per
    n_bad += bad
    kinds[(scoring, flags | ext)] = kinds.get((scoring, flags | ext), 0) + per
    if bad:
        print("MISMATCH", w["scoring"], w["flags"], w["size"], w["x_drop"], w["matrix"], w["gaps"], flush=True)
print(f"worker {wid}: {n_batches} batches, {n_pairs} pairs, {n_bad} mismatching pairs, {len(kinds)} (scoring, flags) combinations")
