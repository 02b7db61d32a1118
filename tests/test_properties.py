"""Size-independent properties of the hot path (mitigation 2) -- checks that do not depend on the
oracle being right, because every reference unit test uses min == max block size and so pins nothing adaptive:

* a block that covers the whole DP matrix must reproduce a plain affine-gap (Gotoh) DP written here in numpy;
* the adaptive aligner can never score above that optimum;
* every CIGAR consumes exactly (query_idx, reference_idx) -- the predicate of examples/verify_trace.rs:8-29 --
  and re-scores, with the same matrix and gaps, to exactly the reported score.

They run against the CPU oracle and against the device source (emulated here; on the GPU in test_gpu_parity.py).
"""
import numpy as np
import pytest

import backend
import ora
import parity
from block_aligner_b200 import api, workloads

P = workloads.params
NEG = -10 ** 9


def gotoh_global(q, r, score, go, ge):
    """Optimal global affine-gap score; a gap of length d costs go + (d - 1) * ge (scores.rs:329-333)."""
    n, m = len(q), len(r)
    D = np.full(m + 1, NEG, dtype=np.int64)
    C = np.full(m + 1, NEG, dtype=np.int64)       # gap along the reference (horizontal)
    D[0] = 0
    for j in range(1, m + 1):
        C[j] = max(C[j - 1] + ge, D[j - 1] + go)
        D[j] = C[j]
    R = np.full(m + 1, NEG, dtype=np.int64)       # gap along the query (vertical), per column
    for i in range(1, n + 1):
        Dp = D.copy()
        R = np.maximum(R + ge, Dp + go)
        D = np.full(m + 1, NEG, dtype=np.int64)
        D[0] = R[0]
        s = score(q[i - 1], r)                    # vector of substitution scores against every r[j]
        diag = Dp[:-1] + s
        c = NEG
        for j in range(1, m + 1):
            c = max(c + ge, D[j - 1] + go)
            D[j] = max(diag[j - 1], c, R[j])
    return int(D[m])


def nuc_scorer(match, mismatch):
    return lambda a, r: np.where(np.frombuffer(r, dtype=np.uint8) == a, match, mismatch).astype(np.int64)


def aa_scorer(mat):
    m = mat.reshape(27, 32).astype(np.int64)
    return lambda a, r: m[a - 65][np.frombuffer(r, dtype=np.uint8) - 65]


def rescore(q, r, runs, score, go, ge):
    """Walk a CIGAR (runs = (len << 4) | op, forward order) -> (consumed q, consumed r, score of the path)."""
    i = j = 0
    total = 0
    for run in runs:
        op, ln = int(run) & 15, int(run) >> 4
        if op in (1, 2, 3):                       # M, =, X
            for t in range(ln):
                total += int(score(q[i + t], r[j + t:j + t + 1])[0])
            i += ln; j += ln
        elif op == 4:                             # I: consumes the query
            total += go + (ln - 1) * ge
            i += ln
        else:                                     # D: consumes the reference
            total += go + (ln - 1) * ge
            j += ln
    return i, j, total


def _pairs(w, n, seed):
    qa, qo, ra, ro = workloads.generate(w["gen"], n, seed=seed, stream=w.get("stream", 0))
    return qa, qo, ra, ro, [(qa[int(qo[k]):int(qo[k + 1])].tobytes(), ra[int(ro[k]):int(ro[k + 1])].tobytes()) for k in range(n)]


DNA = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), x_drop=0, flags=0, stream=31,
           gen=P(alphabet=0, len_dist=0, len_min=40, len_max=220, sub_rate=0.08, ins_rate=0.05, del_rate=0.05,
                 long_indel_mean=1.0, long_indel_len=12.0))
PROT = dict(scoring=api.SCORING_AA, matrix="BLOSUM62", gaps=(-11, -1), x_drop=0, flags=0, stream=32,
            gen=P(alphabet=1, len_dist=0, len_min=30, len_max=200, sub_rate=0.25, ins_rate=0.04, del_rate=0.04))


@pytest.fixture(scope="module")
def env():
    lib = backend.emu_lib()
    return lib, api.Aligner(lib)


def _scorer(lib, w):
    return nuc_scorer(1, -1) if w["scoring"] == api.SCORING_NUC else aa_scorer(lib.builtin_matrix(w["matrix"])[1])


@pytest.mark.parametrize("w", [DNA, PROT], ids=["dna", "protein"])
def test_whole_matrix_block_equals_exact_affine_dp(env, w):
    """min == max >= the longer sequence + 1: the first block is the whole matrix"""
    lib, al = env
    n = 24
    qa, qo, ra, ro, pairs = _pairs(w, n, 5)
    size = (256, 256)
    matrix = workloads.matrix_of(lib, w)
    exp = parity.oracle_batch(w["scoring"], matrix, w["gaps"], size, 0, 0, False, qa, qo, ra, ro)
    got = parity.run_lib(lib, al, w["scoring"], matrix, w["gaps"], size, 0, 0, False, qa, qo, ra, ro)
    sc = _scorer(lib, w)
    for k, (q, r) in enumerate(pairs):
        opt = gotoh_global(q, r, sc, *w["gaps"])
        assert int(exp[0][k][0]) == opt, f"oracle pair {k}"
        assert int(got[0][k][0]) == opt, f"device source pair {k}"


@pytest.mark.parametrize("w", [DNA, PROT], ids=["dna", "protein"])
@pytest.mark.parametrize("size", [(16, 16), (32, 32), (32, 128)])
def test_adaptive_score_never_exceeds_optimum(env, w, size):
    lib, al = env
    qa, qo, ra, ro, pairs = _pairs(w, 16, 9)
    matrix = workloads.matrix_of(lib, w)
    exp = parity.oracle_batch(w["scoring"], matrix, w["gaps"], size, 0, 0, False, qa, qo, ra, ro)
    sc = _scorer(lib, w)
    n_opt = 0
    for k, (q, r) in enumerate(pairs):
        opt = gotoh_global(q, r, sc, *w["gaps"])
        assert int(exp[0][k][0]) <= opt
        n_opt += int(exp[0][k][0]) == opt
    assert n_opt >= len(pairs) // 2      # and it is usually optimal on related sequences


def check_cigars(lib, al, w, n, seed, size, flags, x_drop, scorer):
    """-> number of pairs whose CIGAR is inconsistent, for (oracle, library)"""
    qa, qo, ra, ro, pairs = _pairs(w, n, seed)
    matrix = workloads.matrix_of(lib, w)
    bad = [0, 0]
    outs = [parity.oracle_batch(w["scoring"], matrix, w["gaps"], size, x_drop, flags, True, qa, qo, ra, ro),
            parity.run_lib(lib, al, w["scoring"], matrix, w["gaps"], size, x_drop, flags, True, qa, qo, ra, ro)[:3]]
    for which, (res, cells, cigs) in enumerate(outs):
        for k, (q, r) in enumerate(pairs):
            score, qi, rj = (int(v) for v in res[k])
            ci, cj, total = rescore(q, r, cigs[k], scorer, *w["gaps"])
            if (ci, cj) != (qi, rj) or total != score:
                bad[which] += 1
            # '=' runs must sit on equal bytes and 'X' runs on different ones (cigar_eq, scan_block.rs:1620-1628)
            i = j = 0
            for run in cigs[k]:
                op, ln = int(run) & 15, int(run) >> 4
                if op == 2:
                    assert q[i:i + ln] == r[j:j + ln]
                if op == 3:
                    assert all(q[i + t] != r[j + t] for t in range(ln))
                i += ln if op != 5 else 0
                j += ln if op != 4 else 0
    return bad


@pytest.mark.parametrize("w", [DNA, PROT], ids=["dna", "protein"])
@pytest.mark.parametrize("size,flags,x_drop", [((32, 32), api.TRACE, 0), ((32, 256), api.TRACE, 0),
                                               ((32, 256), api.TRACE | api.XDROP, 40)])
def test_cigar_consumes_end_position_and_rescores(env, w, size, flags, x_drop):
    lib, al = env
    assert check_cigars(lib, al, w, 20, 13, size, flags, x_drop, _scorer(lib, w)) == [0, 0]
