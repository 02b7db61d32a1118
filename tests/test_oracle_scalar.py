"""Two CPU restatements of the reference against each other (SURVEY.md 8c mitigation 1): the literal-intrinsic oracle
(oracle/ba_oracle.cpp, same _mm256_* instructions as src/avx2.rs) and the scalar closed-form one (oracle/ba_scalar.cpp,
one cell at a time). They must agree on the result, the computed cells AND every step of the adaptive state machine
(direction, position, block size, offset, block maximum, border maxima) -- the part of the algorithm that none of the
reference's own unit tests pins (they all use min == max). The quick version runs in the CPU suite; tools/oracle_cross.py
runs the same comparison over >= 10^6 pairs (profiles/r02_oracle_cross.txt)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ora
from block_aligner_b200 import api, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SCA = None


def sca():
    global _SCA
    if _SCA is None:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(os.path.join(ROOT, "oracle", "libba_scalar.so"))
        sz, vp, i32 = C.c_size_t, C.c_void_p, C.c_int
        L.sca_align.argtypes = [vp, sz, vp, sz, i32, vp, i32, i32, sz, sz, i32, i32, vp, vp, vp, vp, vp, sz, vp]
        L.sca_align_profile.argtypes = [vp, sz, vp, vp, vp, vp, sz, i32, sz, sz, i32, i32, vp, vp, vp, vp, vp, sz, vp]
        _SCA = L
    return _SCA


def scalar_align(q, r, kind, matrix, gaps, size, x_drop, flags, profile=None, log=True):
    """q, r: ora.Padded; profile: ora.Profile -> (result tuple, cells, steps list)"""
    L = sca()
    score, qi, ri, cells, n = C.c_int32(), C.c_size_t(), C.c_size_t(), C.c_uint64(), C.c_size_t()
    cap = 1 << 16
    buf = (ora.Step * cap)() if log else None
    if profile is None:
        m = np.ascontiguousarray(matrix, dtype=np.int8)
        e = L.sca_align(q.ptr, len(q), r.ptr, len(r), kind, m.ctypes.data, gaps[0], gaps[1], size[0], size[1], x_drop, flags,
                        C.byref(score), C.byref(qi), C.byref(ri), C.byref(cells), buf, cap, C.byref(n))
    else:
        O = ora.lib()
        e = L.sca_align_profile(q.ptr, len(q), O.ora_profile_pos_aa(profile.h), O.ora_profile_gap_open_C(profile.h),
                                O.ora_profile_gap_close_C(profile.h), O.ora_profile_gap_open_R(profile.h), len(profile),
                                O.ora_profile_get_gap_extend(profile.h), size[0], size[1], x_drop, flags,
                                C.byref(score), C.byref(qi), C.byref(ri), C.byref(cells), buf, cap, C.byref(n))
    assert e == 0
    steps = [(s.dir, s.i, s.j, s.block_size, s.off, s.max, s.right_max, s.down_max) for s in buf[:min(n.value, cap)]] if log else None
    return (score.value, qi.value, ri.value), cells.value, steps


def cross_check(rng, n_pairs, with_steps=True):
    """n_pairs random pairs over random scoring kinds / flags / block ranges -> number of disagreements"""
    O = ora.lib()
    O.ora_profile_pos_aa.restype = C.c_void_p
    for nm in ("open_C", "close_C", "open_R"):
        getattr(O, f"ora_profile_gap_{nm}").restype = C.c_void_p
    P = workloads.params
    bad = done = 0
    while done < n_pairs:
        kind = int(rng.choice([ora.NUC, ora.NUC, ora.AA, ora.BYTE, 3]))
        flags = int(rng.choice([0, ora.XDROP]))
        if kind != 3 and rng.random() < 0.25:
            flags |= int(rng.choice([ora.LOCAL_START, ora.FREE_QUERY_START_GAPS]))
        lo = int(rng.choice([16, 32, 32, 64, 128]))
        hi = int(rng.choice([s for s in (16, 32, 64, 128, 256, 512, 1024) if s >= lo]))
        lmin = int(rng.integers(0, 300))
        lmax = lmin + int(rng.integers(0, 1200))
        rate = float(rng.choice([0.0, 0.03, 0.1, 0.3]))
        gen = P(alphabet=0 if kind in (ora.NUC, ora.BYTE) else 1, len_dist=0, len_min=lmin, len_max=lmax, sub_rate=rate, ins_rate=rate / 2,
                del_rate=rate / 2, long_indel_mean=float(rng.choice([0.0, 2.0])), long_indel_len=float(rng.choice([20.0, 100.0])),
                suffix_len=int(rng.choice([0, 80])), big_indel_prob=float(rng.choice([0.0, 0.5])), big_indel_min=40, big_indel_max=300)
        k = int(min(16, n_pairs - done))
        qa, qo, ra, ro = workloads.generate(gen, k, seed=int(rng.integers(1 << 30)), stream=int(rng.integers(1, 999)))
        x_drop = int(rng.choice([5, 50, 400]))
        if kind == ora.NUC:
            matrix = ora.nw1() if rng.random() < 0.5 else ora.nuc_matrix(int(rng.integers(1, 6)), -int(rng.integers(1, 7)))
            e = -int(rng.integers(1, 4))
            gaps = (e - int(rng.integers(1, 9)), e)
        elif kind == ora.AA:
            matrix, gaps = ora.builtin(str(rng.choice(["BLOSUM62", "PAM250"]))), (-11, -1)
        elif kind == ora.BYTE:
            matrix, gaps = np.array([1, -1], dtype=np.int8), (-2, -1)
        else:
            matrix, gaps = None, None
        for t in range(k):
            qb, rb = qa[int(qo[t]):int(qo[t + 1])].tobytes(), ra[int(ro[t]):int(ro[t + 1])].tobytes()
            pk = ora.AA if kind == 3 else kind
            q = ora.Padded(pk, qb, hi)
            blk = ora.Block(len(qb), len(rb), hi, flags)
            blk.enable_step_log(with_steps)
            if kind == 3:
                prof = ora.Profile.new(len(rb), hi, -1)
                if len(rb):
                    sc = workloads.pssm_scores(np.asarray(ora.builtin("BLOSUM62"), dtype=np.int8), rb, rng)
                    assert O.ora_profile_set_all(prof.h, workloads.MAP20, 20, sc.ctypes.data, sc.size, 0, 0, 0) == 0
                for i in range(1, len(rb) + 1):
                    prof.set_gap_open_C(i, int(rng.integers(-12, -5))); prof.set_gap_close_C(i, int(rng.integers(-2, 1))); prof.set_gap_open_R(i, int(rng.integers(-12, -5)))
                exp = blk.align_profile(q, prof, (lo, hi), x_drop)
                got = scalar_align(q, None, 3, None, None, (lo, hi), x_drop, flags, profile=prof, log=with_steps)
            else:
                r = ora.Padded(pk, rb, hi)
                exp = blk.align(q, r, kind, matrix, gaps, (lo, hi), x_drop)
                got = scalar_align(q, r, kind, matrix, gaps, (lo, hi), x_drop, flags, log=with_steps)
            ok = tuple(exp) == got[0] and blk.cells() == got[1]
            if ok and with_steps:
                ok = blk.logged_steps() == got[2]
            if not ok:
                bad += 1
                if bad <= 3:
                    print("DISAGREE", kind, flags, (lo, hi), gaps, x_drop, len(qb), len(rb), tuple(exp), got[0], blk.cells(), got[1])
                    if with_steps:
                        a, b = blk.logged_steps(), got[2]
                        for s, (u, v) in enumerate(zip(a, b)):
                            if u != v:
                                print("  first differing step", s, u, v)
                                break
                        else:
                            print("  step counts", len(a), len(b))
            done += 1
    return bad


def test_scalar_restatement_matches_reference_unit_vectors():
    """a few of the reference's own assertions (src/scan_block.rs:1914-2049) through the scalar restatement"""
    b62, nw1 = ora.builtin("BLOSUM62"), ora.nw1()
    for q, r, s in [(b"AARA", b"AAAA", 11), (b"AARAAAA", b"AAAAAAAA", 12), (b"RRRR", b"AAAA", -4), (b"AAA", b"AAAA", 1), (b"", b"AAAA", -14)]:
        got = scalar_align(ora.Padded(ora.AA, q, 16), ora.Padded(ora.AA, r, 16), ora.AA, b62, (-11, -1), (16, 16), 0, 0)
        assert got[0] == (s, len(q), len(r))
    got = scalar_align(ora.Padded(ora.NUC, b"TTTTTTTTAAAAAAATTTTTTTTT", 16), ora.Padded(ora.NUC, b"TTAAAAAAATTTTTTTTTTTT", 16), ora.NUC, nw1,
                       (-2, -1), (16, 16), 0, 0)
    assert got[0] == (7, 24, 21)
    got = scalar_align(ora.Padded(ora.AA, b"A" * 44, 16), ora.Padded(ora.AA, b"A" * 15 + b"R" * 16 + b"A" * 13, 16), ora.AA, b62, (-11, -1),
                       (16, 16), 1, ora.XDROP)
    assert got[0] == (60, 15, 15)


def test_two_restatements_agree_step_by_step():
    assert cross_check(np.random.default_rng(2024), 1500) == 0
