#!/usr/bin/env python3
"""Cross-check of the two CPU restatements (oracle/ba_oracle.cpp, literal AVX2 intrinsics, vs oracle/ba_scalar.cpp, scalar
closed form) over many random pairs: result, computed cells and every step of the state machine must agree.
usage: tools/oracle_cross.py <pairs per worker> <workers>        -> one line per worker + a total"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def work(args):
    wid, n = args
    import numpy as np
    import test_oracle_scalar as t
    bad = t.cross_check(np.random.default_rng(777 + wid), n)
    return wid, n, bad


if __name__ == "__main__":
    n, workers = int(sys.argv[1]), int(sys.argv[2])
    import test_oracle_scalar as t
    t.sca()        # build once
    tot = bad = 0
    with ProcessPoolExecutor(workers) as ex:
        for wid, k, b in ex.map(work, [(w, n) for w in range(workers)]):
            print(f"worker {wid}: {k} pairs, {b} disagreements", flush=True)
            tot += k
            bad += b
    print(f"total: {tot} pairs (scoring kinds nuc/aa/byte/profile, flags 0 / X_DROP / LOCAL_START / FREE_QUERY_START_GAPS, blocks 16..1024), "
          f"{bad} disagreements on result, cells or any step")
