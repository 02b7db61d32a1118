#!/usr/bin/env python3
"""Randomised soak against the oracle: many batches over random scoring kinds, flag combinations (TRACE, X_DROP,
LOCAL_START, FREE_QUERY_START_GAPS, FREE_QUERY_END_GAPS), block ranges, length ranges and error profiles. Every pair is
compared on score, end indices, CIGAR and the path-determined cell count (tests/parity.py).

usage: tools/soak.py <cuda|emu> <minutes> <worker id> [pairs per batch]      -> one summary line on stdout
  cuda: the product library on the GPU (large batches; run under gpurun)
  emu:  the device source compiled for the CPU with the fiber emulator (small batches; several workers in parallel)
Test infrastructure (uses tests/ and oracle/).
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
from test_emu_parity import _short_query_batch  # noqa: E402

which, minutes, wid = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
per = int(sys.argv[4]) if len(sys.argv) > 4 else (8 if which == "emu" else 512)
rng = np.random.default_rng(1000 + wid)
lib = backend.emu_lib() if which == "emu" else backend.cuda_lib()
al = api.Aligner(lib)
P = workloads.params
t_end = time.time() + 60 * minutes
n_pairs = n_bad = n_batches = 0
kinds = {}
while time.time() < t_end:
    scoring = int(rng.choice([api.SCORING_NUC, api.SCORING_NUC, api.SCORING_AA, api.SCORING_BYTE, api.SCORING_PROFILE]))
    flags = int(rng.choice([0, api.XDROP, api.TRACE, api.TRACE | api.XDROP]))
    ext = 0
    if scoring != api.SCORING_PROFILE and rng.random() < 0.25:
        ext = int(rng.choice([api.LOCAL_START, api.FREE_QUERY_START_GAPS, api.FREE_QUERY_END_GAPS]))
    lo = int(rng.choice([16, 32, 32, 64, 128]))
    hi = int(rng.choice([s for s in (16, 32, 64, 128, 256, 512, 1024, 2048) if s >= lo]))
    seed = int(rng.integers(1 << 30))
    try:
        if ext == api.FREE_QUERY_END_GAPS:
            # short queries (the mode needs min block > |q|, no X_DROP), references up to a few blocks long
            flags &= ~api.XDROP
            if rng.random() < 0.3:
                flags |= int(rng.choice([api.LOCAL_START, api.FREE_QUERY_START_GAPS]))
            lo = int(rng.choice([32, 64, 128, 256]))
            hi = int(rng.choice([s for s in (64, 256, 512, 1024, 2048) if s >= lo]))
            fl = flags | ext
            kind = int(rng.choice([api.SCORING_NUC, api.SCORING_AA]))
            alpha = b"ACGT" if kind == api.SCORING_NUC else b"ACDEFGHIKLMNPQRSTVWY"
            m = lib.builtin_matrix("NW1" if kind == api.SCORING_NUC else "BLOSUM62")[1]
            gaps = (-2, -1) if kind == api.SCORING_NUC else (-11, -1)
            nq = min(per, 64)
            qa, qo, ra, ro = _short_query_batch(nq, lo - 1, 40, int(rng.choice([300, 900, 3000])), seed=seed, alphabet=alpha)
            got = parity.run_lib(lib, al, kind, m, gaps, (lo, hi), 0, fl, bool(fl & api.TRACE), qa, qo, ra, ro)
            exp = parity.oracle_batch(kind, m, gaps, (lo, hi), 0, fl, bool(fl & api.TRACE), qa, qo, ra, ro)
            bad = parity.compare("soak-fqe", got, exp)
            scoring, npairs = kind, nq
            if bad:
                print("MISMATCH fqe", kind, fl, (lo, hi), seed, flush=True)
        else:
            lmin = int(rng.integers(1, 400))
            lmax = lmin + int(rng.integers(0, 1500 if which == "emu" else 6000))
            rate = float(rng.choice([0.0, 0.02, 0.05, 0.12, 0.3]))
            gen = P(alphabet=0 if scoring in (api.SCORING_NUC, api.SCORING_BYTE) else 1, len_dist=0, len_min=lmin, len_max=lmax,
                    sub_rate=rate, ins_rate=rate / 2, del_rate=rate / 2, long_indel_mean=float(rng.choice([0.0, 1.0, 3.0])),
                    long_indel_len=float(rng.choice([10.0, 60.0, 200.0])), suffix_len=int(rng.choice([0, 50, 300])),
                    big_indel_prob=float(rng.choice([0.0, 0.5])), big_indel_min=50, big_indel_max=400)
            if scoring == api.SCORING_NUC:
                matrix, gaps = (("NW1", (-2, -1)) if rng.random() < 0.5 else ((int(rng.integers(1, 6)), -int(rng.integers(1, 7))), None))
                if gaps is None:
                    e = -int(rng.integers(1, 4))
                    gaps = (e - int(rng.integers(1, 9)), e)
            elif scoring == api.SCORING_AA:
                matrix, gaps = str(rng.choice(["BLOSUM62", "BLOSUM45", "PAM250"])), (-11, -1)
            elif scoring == api.SCORING_BYTE:
                matrix, gaps = (1, -1), (-2, -1)
            else:
                matrix, gaps = None, None
            w = dict(scoring=scoring, matrix=matrix, gaps=gaps, size=(lo, hi), x_drop=int(rng.choice([0, 10, 50, 400])), flags=flags | ext,
                     stream=int(rng.integers(1, 1000)), gen=gen)
            npairs = per if scoring != api.SCORING_PROFILE else min(per, 64)     # profiles are built position by position in Python
            if scoring == api.SCORING_BYTE:
                # ByteMatrix pads with byte 0, which *matches* the other sequence's padding, so an X-drop alignment can end
                # past both sequences; the reference's traceback then panics ("Traceback cigar end position must be in
                # bounds!", scan_block.rs:1483) -- there is no CIGAR to compare. Keep TRACE, drop X_DROP for this kind.
                if (w["flags"] & api.TRACE) and (w["flags"] & api.XDROP):
                    w["flags"] &= ~api.XDROP
                    flags &= ~api.XDROP
                qa, qo, ra, ro = workloads.generate(gen, npairs, seed=seed, stream=w["stream"])
                m = np.array([1, -1], dtype=np.int8)
                got = parity.run_lib(lib, al, scoring, m, gaps, w["size"], w["x_drop"], w["flags"], bool(flags & api.TRACE), qa, qo, ra, ro)
                exp = parity.oracle_batch(scoring, m, gaps, w["size"], w["x_drop"], w["flags"], bool(flags & api.TRACE), qa, qo, ra, ro)
                bad = parity.compare("soak-byte", got, exp)
            else:
                bad = parity.check_workload(lib, al, w, npairs, seed=seed)
            if bad:
                print("MISMATCH", w["scoring"], w["flags"], w["size"], w["x_drop"], w["matrix"], w["gaps"], seed, w["stream"], flush=True)
    except api.BlockAlignerError:      # argument combinations the reference rejects too (e.g. open >= extend)
        continue
    n_batches += 1
    n_pairs += npairs
    n_bad += bad
    kinds[(scoring, flags | ext)] = kinds.get((scoring, flags | ext), 0) + npairs
print(f"{which} worker {wid}: {n_batches} batches, {n_pairs} pairs, {n_bad} mismatching pairs, {len(kinds)} (scoring, flags) combinations")
