"""Parity harness: run a batch through the C-ABI library (CUDA product or the emulated test build)
and through the CPU oracle on identical inputs, compare score, end indices, CIGAR, and the
path-determined work counters (cells, steps). Test infrastructure.
"""
import ctypes as C

import numpy as np

import ora
from block_aligner_b200 import api, workloads


def pad_arena(scoring, arena, off, pad):
    """PaddedBytes::from_bytes for every sequence of an arena (done by the oracle library in C).
    Returns (padded arena, padded offsets uint64[n], lengths uint32[n])."""
    n = len(off) - 1
    kind = {api.SCORING_NUC: ora.NUC, api.SCORING_AA: ora.AA, api.SCORING_PROFILE: ora.AA, api.SCORING_BYTE: ora.BYTE}[scoring]
    off = np.ascontiguousarray(off, dtype=np.uint64)
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    total = int(off[-1] - off[0]) + n * (1 + pad)
    out = np.zeros(total + 64, dtype=np.uint8)
    poff = np.zeros(max(n, 1), dtype=np.uint64)
    lens = np.zeros(max(n, 1), dtype=np.uint32)
    e = ora.lib().ora_pad_batch(kind, arena.ctypes.data, off.ctypes.data, n, pad, out.ctypes.data, poff.ctypes.data, lens.ctypes.data)
    if e:
        raise ValueError(f"ora_pad_batch error {e}")
    return out, poff[:n], lens[:n]


_RES_DT = np.dtype([("score", np.int32), ("query_idx", np.uint64), ("reference_idx", np.uint64)], align=True)
_OPLEN_DT = np.dtype([("op", np.uint8), ("len", np.uint64)], align=True)


def oracle_batch(scoring, matrix, gaps, size, x_drop, flags, cigar_eq, qa, qo, ra, ro, profiles=None, threads=None,
                 cigars_as_arrays=True):
    """-> results (n x 3 int64), cells (uint64[n]), cigars (list of run arrays (len << 4 | op, uint64) or None)"""
    L = ora.lib()
    n = len(qo) - 1
    pad = max(size[1], 16) + 32
    pq, pqo, ql = pad_arena(scoring, qa, qo, pad)
    if profiles is None:
        pr, pro, rl = pad_arena(scoring, ra, ro, pad)
        arena = np.concatenate([pq, pr])
        pro = pro + np.uint64(pq.size)
    else:
        arena, pro, rl = pq, np.zeros(n, dtype=np.uint64), np.array([len(p) for p in profiles], dtype=np.uint32)
    b = ora.Batch()
    b.n = n
    b.arena = arena.ctypes.data
    b.q_off, b.q_len = pqo.ctypes.data, ql.ctypes.data
    b.r_off, b.r_len = pro.ctypes.data, rl.ctypes.data
    parr = None
    if profiles is not None:
        parr = (C.c_void_p * n)(*[p.h for p in profiles])
        b.profiles = C.cast(parr, C.c_void_p)
    else:
        b.profiles = None
    mat = None
    if matrix is not None:
        mat = np.ascontiguousarray(matrix, dtype=np.int8)
        b.matrix = mat.ctypes.data
    b.matrix_kind = {api.SCORING_NUC: ora.NUC, api.SCORING_AA: ora.AA, api.SCORING_BYTE: ora.BYTE,
                     api.SCORING_PROFILE: ora.AA}[scoring]
    if gaps:
        b.gap_open, b.gap_extend = gaps
    b.min_size, b.max_size = size
    b.x_drop, b.flags, b.cigar_eq = x_drop, flags, int(bool(cigar_eq))
    assert _RES_DT.itemsize == C.sizeof(ora.Result) and _OPLEN_DT.itemsize == C.sizeof(ora.OpLen)
    out = np.zeros(max(n, 1), dtype=_RES_DT)
    cells = np.zeros(max(n, 1), dtype=np.uint64)
    trace = bool(flags & api.TRACE)
    coff = clen = carena = None
    if trace:
        caps = ql.astype(np.uint64) + rl.astype(np.uint64) + 5
        coff = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(caps, out=coff[1:])
        carena = np.zeros(int(coff[-1] + 1), dtype=_OPLEN_DT)
        clen = np.zeros(max(n, 1), dtype=np.uint32)
    threads = threads or ora.lib().ora_hw_threads()
    e = L.ora_batch_align(C.byref(b), threads, out.ctypes.data, cells.ctypes.data, carena.ctypes.data if trace else None,
                          coff.ctypes.data if trace else None, clen.ctypes.data if trace else None)
    if e:
        raise ValueError(f"oracle batch error {e}")
    res = np.stack([out["score"][:n].astype(np.int64), out["query_idx"][:n].astype(np.int64),
                    out["reference_idx"][:n].astype(np.int64)], axis=1).reshape(n, 3)
    cigs = None
    if trace:
        packed = (carena["len"] << np.uint64(4)) | carena["op"].astype(np.uint64)
        cigs = [packed[int(coff[k]):int(coff[k]) + int(clen[k])] for k in range(n)]
    return res, cells[:n], cigs


def make_ora_profiles(ra, ro, block_size, gap_open, gap_extend, seed):
    """The C4 profiles of workloads.make_pssm_batch / make_lib_profiles for the oracle: same numbers, same rng order
    (PSSM = BLOSUM62 row of the consensus + noise; gaps for positions 1..=len, position 0 keeps the -128 defaults)."""
    b62 = np.asarray(ora.builtin("BLOSUM62"), dtype=np.int8)
    rng = np.random.default_rng(seed)
    L = ora.lib()
    out = []
    for k in range(len(ro) - 1):
        cons = ra[int(ro[k]):int(ro[k + 1])].tobytes()
        o = ora.Profile.new(len(cons), block_size, gap_extend)
        if len(cons):
            sc = workloads.pssm_scores(b62, cons, rng)
            assert L.ora_profile_set_all(o.h, workloads.MAP20, 20, sc.ctypes.data, sc.size, 0, 0, 0) == 0
        assert L.ora_profile_set_all_gap_open_C(o.h, gap_open) == 0
        assert L.ora_profile_set_all_gap_close_C(o.h, 0) == 0
        assert L.ora_profile_set_all_gap_open_R(o.h, gap_open) == 0
        o.set_gap_open_C(0, -128); o.set_gap_close_C(0, -128); o.set_gap_open_R(0, -128)
        out.append(o)
    return out


def make_profiles(lib, b62, ra, ro, block_size, gap_open, gap_extend, seed):
    """C4 profiles: the reference side of each pair is the consensus; PSSM = BLOSUM62 row + noise;
    gaps set for positions 1..=len like examples/pssm_bench.rs:67-83 (position 0 keeps the -128 defaults).
    Built twice from the same numbers: for the library (block_*_aaprofile) and for the oracle."""
    n = len(ro) - 1
    rng = np.random.default_rng(seed)
    lib_profs, ora_profs = [], []
    for k in range(n):
        cons = ra[int(ro[k]):int(ro[k + 1])].tobytes()
        sc = workloads.pssm_scores(b62, cons, rng)
        p = api.AAProfile(lib, len(cons), block_size, gap_extend)
        o = ora.Profile.new(len(cons), block_size, gap_extend)
        if len(cons):
            p.set_all(workloads.MAP20, sc)
            assert ora.lib().ora_profile_set_all(o.h, workloads.MAP20, 20, sc.ctypes.data, sc.size, 0, 0, 0) == 0
        for i in range(1, len(cons) + 1):
            p.set_gap_open_C(i, gap_open); p.set_gap_close_C(i, 0); p.set_gap_open_R(i, gap_open)
            o.set_gap_open_C(i, gap_open); o.set_gap_close_C(i, 0); o.set_gap_open_R(i, gap_open)
        lib_profs.append(p)
        ora_profs.append(o)
    return lib_profs, ora_profs


def run_lib(lib, al, scoring, matrix, gaps, size, x_drop, flags, cigar_eq, qa, qo, ra, ro, profiles=None):
    cfg = al.config(scoring, matrix, gaps, size, x_drop, flags, cigar_eq)
    b = al.upload(cfg, qa, qo, ra, ro, profiles)
    try:
        st = b.run()
        r = b.download()
        n = len(qo) - 1
        res = np.stack([r["score"].astype(np.int64), r["query_idx"].astype(np.int64), r["reference_idx"].astype(np.int64)], axis=1).reshape(n, 3)
        cells = np.array([b.pair_stats(k)[0] for k in range(n)], dtype=np.uint64)
        cigs = [b.cigar_runs(k).astype(np.uint64) for k in range(n)] if flags & api.TRACE else None
        return res, cells, cigs, st
    finally:
        b.free()


def compare(tag, got, exp, verbose=True, limit=5):
    """got/exp = (res, cells, cigs). Returns the number of mismatching pairs."""
    res_g, cells_g, cig_g = got[:3]
    res_e, cells_e, cig_e = exp
    n = len(res_e)
    bad = []
    for k in range(n):
        ok = (res_g[k] == res_e[k]).all() and int(cells_g[k]) == int(cells_e[k])
        if ok and cig_e is not None:
            ok = len(cig_g[k]) == len(cig_e[k]) and (cig_g[k] == cig_e[k]).all()
        if not ok:
            bad.append(k)
    if bad and verbose:
        for k in bad[:limit]:
            msg = f"[{tag}] pair {k}: got {res_g[k].tolist()} cells {int(cells_g[k])}  expected {res_e[k].tolist()} cells {int(cells_e[k])}"
            if cig_e is not None:
                msg += f"\n   cigar got {api.runs_to_string(cig_g[k])[:120]}\n   cigar exp {api.runs_to_string(cig_e[k])[:120]}"
            print(msg)
        print(f"[{tag}] {len(bad)} / {n} pairs differ")
    return len(bad)


def check_workload(lib, al, w, n, first=0, seed=1234, with_trace=None, size=None, flags=None, verbose=True):
    """Generate n pairs of workload dict `w`, run library and oracle, return the mismatch count."""
    scoring = w["scoring"]
    flags = w["flags"] if flags is None else flags
    if with_trace is True:
        flags |= api.TRACE
    elif with_trace is False:
        flags &= ~api.TRACE
    size = size or w["size"]
    cigar_eq = bool(flags & api.TRACE) and scoring != api.SCORING_PROFILE
    qa, qo, ra, ro = workloads.generate(w["gen"], n, first=first, seed=seed, stream=w.get("stream", 0))
    matrix = workloads.matrix_of(lib, w)
    lp = op = None
    if scoring == api.SCORING_PROFILE:
        b62 = lib.builtin_matrix("BLOSUM62")[1]
        lp, op = make_profiles(lib, b62, ra, ro, size[1], -10, -1, seed + first)
    got = run_lib(lib, al, scoring, matrix, w["gaps"], size, w["x_drop"], flags, cigar_eq, qa, qo, ra, ro, lp)
    exp = oracle_batch(scoring, matrix, w["gaps"], size, w["x_drop"], flags, cigar_eq, qa, qo, ra, ro, op)
    return compare(f"{scoring}/{flags}/{size}", got, exp, verbose)


def check_pssm(lib, al, n, seed, rev, shifts, uniform_gaps, size=(32, 128), flags=api.XDROP):
    """The same numbers go to the oracle through Profile::new + set_all[_rev] + gap setters (scores.rs:532-580)."""
    w = workloads.WORKLOADS["C4_seq_to_profile_xdrop"]
    qa, qo, ra, ro = workloads.generate(w["gen"], n, seed=seed, stream=4)
    b62 = lib.builtin_matrix("BLOSUM62")[1]
    rng = np.random.default_rng(seed)
    ls, rs = shifts
    scores, soff, ops = [], [0], []
    go, gc, gr = [], [], []
    for k in range(n):
        cons = ra[int(ro[k]):int(ro[k + 1])].tobytes()
        sc = workloads.pssm_scores(b62, cons, rng)
        scores.append(sc.reshape(-1))
        soff.append(soff[-1] + sc.size)
        o = ora.Profile.new(len(cons), size[1], -1)
        if len(cons):
            assert ora.lib().ora_profile_set_all(o.h, workloads.MAP20, 20, sc.ctypes.data, sc.size, ls, rs, int(rev)) == 0
        g = rng.integers(-12, -5, size=(3, len(cons) + 1)).astype(np.int8)
        g[1] = rng.integers(-2, 1, size=len(cons) + 1)
        if uniform_gaps:
            g[0, :] = -9; g[1, :] = -1; g[2, :] = -7
        for i in range(len(cons) + 1):
            o.set_gap_open_C(i, int(g[0, i])); o.set_gap_close_C(i, int(g[1, i])); o.set_gap_open_R(i, int(g[2, i]))
        go.append(g[0]); gc.append(g[1]); gr.append(g[2])
        ops.append(o)
    scores = np.concatenate(scores) if scores else np.zeros(0, dtype=np.int8)
    if uniform_gaps:
        pb = api.PssmBatch(workloads.MAP20, scores, soff, -1, all_gaps=(-9, -1, -7), left_shift=ls, right_shift=rs, rev=rev)
    else:
        pb = api.PssmBatch(workloads.MAP20, scores, soff, -1, gaps=(np.concatenate(go), np.concatenate(gc), np.concatenate(gr)),
                           left_shift=ls, right_shift=rs, rev=rev)
    got = run_lib(lib, al, api.SCORING_PROFILE, None, None, size, w["x_drop"], flags, False, qa, qo, ra, ro, pb)
    exp = oracle_batch(api.SCORING_PROFILE, None, None, size, w["x_drop"], flags, False, qa, qo, ra, ro, ops)
    return compare("pssm", got, exp)


def check_reversed(lib, al, n, rev, seed=5):
    """Library with BA_REV_* flags on the original inputs vs the oracle on inputs reversed on the host."""
    w = workloads.WORKLOADS["C2_nanopore_xdrop_10k"]
    gen = workloads.params(alphabet=0, len_dist=0, len_min=200, len_max=1500, sub_rate=0.05, ins_rate=0.04, del_rate=0.04,
                           long_indel_mean=1.0, long_indel_len=40.0, suffix_len=100)
    qa, qo, ra, ro = workloads.generate(gen, n, seed=seed, stream=2)

    def reversed_arena(a, off):
        out = a.copy()
        for k in range(len(off) - 1):
            out[int(off[k]):int(off[k + 1])] = a[int(off[k]):int(off[k + 1])][::-1]
        return out
    qh = reversed_arena(qa, qo) if rev & api.REV_QUERY else qa
    rh = reversed_arena(ra, ro) if rev & api.REV_REFERENCE else ra
    m = workloads.matrix_of(lib, w)
    flags = api.XDROP | api.TRACE
    got = run_lib(lib, al, w["scoring"], m, w["gaps"], (32, 256), 50, flags | rev, True, qa, qo, ra, ro)
    exp = oracle_batch(w["scoring"], m, w["gaps"], (32, 256), 50, flags, True, qh, qo, rh, ro)
    return compare(f"rev{rev}", got, exp)


# ---- the one-call entry points (ba_align_batch*), as a C caller uses them ---------------------------------------
ABI_RES_DT = np.dtype([("score", np.int32), ("q", np.uint64), ("r", np.uint64)], align=True)


def abi_align_batch(lib, al, cfg, qa, qo, ra, ro):
    """ba_align_batch -> (results n x 3 int64, BaStats)"""
    n = len(qo) - 1
    out = np.zeros(max(n, 1), dtype=ABI_RES_DT)
    st = api.BaStats()
    lib.check(lib.L.ba_align_batch(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data,
                                   out.ctypes.data, C.byref(st)))
    return np.stack([out["score"][:n].astype(np.int64), out["q"][:n].astype(np.int64), out["r"][:n].astype(np.int64)], axis=1), st


def abi_align_batch_cigar(lib, al, cfg, qa, qo, ra, ro):
    """ba_align_batch_cigar -> (results, list of run arrays, BaStats)"""
    n = len(qo) - 1
    out = np.zeros(max(n, 1), dtype=ABI_RES_DT)
    cap = int(qo[-1] - qo[0] + ro[-1] - ro[0]) + 5 * n + 8
    runs = np.zeros(cap, dtype=np.uint32)
    off = np.zeros(max(n, 1), dtype=np.uint64)
    ln = np.zeros(max(n, 1), dtype=np.uint32)
    used = C.c_size_t()
    st = api.BaStats()
    lib.check(lib.L.ba_align_batch_cigar(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data,
                                         out.ctypes.data, runs.ctypes.data, cap, off.ctypes.data, ln.ctypes.data,
                                         C.byref(used), C.byref(st)))
    assert used.value == int(ln[:n].sum())
    cigs = [runs[int(off[k]):int(off[k]) + int(ln[k])].astype(np.uint64) for k in range(n)]
    return np.stack([out["score"][:n].astype(np.int64), out["q"][:n].astype(np.int64), out["r"][:n].astype(np.int64)], axis=1), cigs, st


def compare_abi(tag, res, cigs, st, exp):
    """results (+ CIGARs) of a one-call entry point against oracle_batch's tuple; also the cell total of BaStats"""
    res_e, cells_e, cig_e = exp
    bad = int((res != res_e).any(axis=1).sum())
    if cigs is not None:
        bad += sum(1 for k in range(len(res_e)) if len(cigs[k]) != len(cig_e[k]) or not (cigs[k] == cig_e[k]).all())
    if st is not None and int(st.cells) != int(cells_e.sum()):
        print(f"[{tag}] cells {int(st.cells)} != oracle {int(cells_e.sum())}")
        bad += 1
    if bad:
        k = int(np.argmax((res != res_e).any(axis=1)))
        print(f"[{tag}] {bad} mismatches; first differing result: pair {k} got {res[k].tolist()} expected {res_e[k].tolist()}")
    return bad
