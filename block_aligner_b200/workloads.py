"""The five BASELINE.json workloads as seeded synthetic batches.

Parameters with a citation come from the reference's benches; everything else (length
distributions, error mix, X-drop thresholds where the reference has no such benchmark) is our
choice, frozen here so that the GPU and the CPU baseline always see byte-identical inputs.
The generator itself is block_aligner_b200/csrc/ba_gen.cpp (splitmix64-seeded xoshiro256** per pair).
"""
import ctypes as C
import os

import numpy as np

from . import api

_HERE = os.path.dirname(os.path.abspath(__file__))
_GEN = None


class GenParams(C.Structure):
    _fields_ = [("alphabet", C.c_int32), ("len_dist", C.c_int32), ("len_min", C.c_uint32), ("len_max", C.c_uint32),
                ("len_median", C.c_double), ("len_sigma", C.c_double), ("k_edits", C.c_int32),
                ("sub_rate", C.c_double), ("ins_rate", C.c_double), ("del_rate", C.c_double),
                ("id_min", C.c_double), ("id_max", C.c_double), ("indel_max", C.c_double),
                ("long_indel_mean", C.c_double), ("long_indel_len", C.c_double), ("big_indel_prob", C.c_double),
                ("big_indel_min", C.c_uint32), ("big_indel_max", C.c_uint32), ("suffix_len", C.c_uint32)]


def _gen():
    global _GEN
    if _GEN is None:
        path = os.path.join(_HERE, "libba_gen.so")
        if not os.path.exists(path):
            raise api.BlockAlignerError(f"{path} not found: run __graft_entry__.build()")
        L = C.CDLL(path)
        L.ba_gen_lengths.argtypes = [C.POINTER(GenParams), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                     C.c_void_p, C.c_int]
        L.ba_gen_fill.argtypes = [C.POINTER(GenParams), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _GEN = L
    return _GEN


def params(**kw):
    p = GenParams()
    p.k_edits = -1
    for k, v in kw.items():
        setattr(p, k, v)
    return p


# name -> dict(gen=GenParams, n=pairs at full size, scoring, matrix (name or (match, mismatch)), gaps, size, x_drop,
#              flags, cigar_eq, stream)
WORKLOADS = {
    # C1 benches/rand_scan.rs:96-109 (NW1, gaps -2/-1, K = len/10 edits) with the fixed block of :63-76
    "C1_rand_scan_dna1k": dict(
        gen=params(alphabet=0, len_dist=0, len_min=1000, len_max=1000, k_edits=100),
        n=1000, scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=(32, 32), x_drop=0, flags=0, stream=1),
    # C2 examples/nanopore_bench.rs:97-104 (NW1, -2/-1, x_drop 50) and :45-48 (unrelated 500-base tails)
    "C2_nanopore_xdrop_10k": dict(
        gen=params(alphabet=0, len_dist=0, len_min=8000, len_max=12000, sub_rate=0.04, ins_rate=0.04, del_rate=0.04,
                   long_indel_mean=1.0, long_indel_len=60.0, suffix_len=500),
        n=100_000, scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=(32, 256), x_drop=50, flags=api.XDROP,
        stream=2),
    # C3 examples/uc_bench.rs:79-104 (BLOSUM62, -11/-1, 32..=256); lengths after the Uniclust30 statistics
    "C3_uniclust_protein_global": dict(
        gen=params(alphabet=1, len_dist=1, len_min=25, len_max=2000, len_median=300.0, len_sigma=0.55,
                   id_min=0.30, id_max=1.0, indel_max=0.10),
        n=1_000_000, scoring=api.SCORING_AA, matrix="BLOSUM62", gaps=(-11, -1), size=(32, 256), x_drop=0, flags=0,
        stream=3),
    # C4 sequence-to-profile (examples/pssm_bench.rs:45-117): profiles are built from the *reference* side of
    # these pairs (consensus) by workloads.make_profiles; queries are the mutated side
    "C4_seq_to_profile_xdrop": dict(
        gen=params(alphabet=1, len_dist=0, len_min=100, len_max=300, sub_rate=0.30, ins_rate=0.02, del_rate=0.02),
        n=100_000, scoring=api.SCORING_PROFILE, matrix=None, gaps=None, size=(32, 256), x_drop=100, flags=api.XDROP,
        stream=4),
    # C5 examples/nanopore_bench_global.rs:150-166 (NucMatrix 2/-4, gaps -6/-2, trace + cigar), X-drop variant
    "C5_longread_trace_50k": dict(
        gen=params(alphabet=0, len_dist=0, len_min=40000, len_max=50000, sub_rate=0.04, ins_rate=0.04, del_rate=0.04,
                   long_indel_mean=1.0, long_indel_len=60.0, big_indel_prob=0.10, big_indel_min=1000, big_indel_max=3000,
                   suffix_len=500),
        n=10_000, scoring=api.SCORING_NUC, matrix=(2, -4), gaps=(-6, -2), size=(64, 2048), x_drop=400,
        flags=api.XDROP | api.TRACE, cigar_eq=True, stream=5),
}


def generate(gen, n, first=0, seed=1234, stream=0, threads=None):
    """-> q_arena (uint8), q_off (uint64[n+1]), r_arena, r_off"""
    L = _gen()
    threads = threads or min(32, os.cpu_count() or 1)
    ql = np.zeros(n, dtype=np.uint32)
    rl = np.zeros(n, dtype=np.uint32)
    L.ba_gen_lengths(C.byref(gen), seed, stream, first, n, ql.ctypes.data, rl.ctypes.data, threads)
    qo = np.zeros(n + 1, dtype=np.uint64)
    ro = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(ql, out=qo[1:], dtype=np.uint64)
    np.cumsum(rl, out=ro[1:], dtype=np.uint64)
    qa = np.zeros(int(qo[-1]), dtype=np.uint8)
    ra = np.zeros(int(ro[-1]), dtype=np.uint8)
    L.ba_gen_fill(C.byref(gen), seed, stream, first, n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data, threads)
    return qa, qo, ra, ro


def matrix_of(lib, w):
    m = w["matrix"]
    if m is None:
        return None
    if isinstance(m, str):
        return lib.builtin_matrix(m)[1]
    return api.nuc_matrix(*m) if w["scoring"] == api.SCORING_NUC else api.aa_matrix_simple(*m)


MAP20 = b"ACDEFGHIKLMNPQRSTVWY"   # examples/pssm_bench.rs:43


def pssm_scores(blosum62, consensus, rng):
    """PSSM rows for a consensus string: BLOSUM62 row of the residue + integer noise in [-1, 1], clipped to i8,
    20 standard residues in MAP20 order."""
    cons = np.frombuffer(bytes(consensus), dtype=np.uint8).astype(np.int64) - 65
    cols = np.frombuffer(MAP20, dtype=np.uint8).astype(np.int64) - 65
    base = blosum62.reshape(27, 32)[cons][:, cols].astype(np.int16)
    noise = rng.integers(-1, 2, size=base.shape, dtype=np.int16)
    return np.clip(base + noise, -128, 127).astype(np.int8)


def make_lib_profiles(lib, ra, ro, block_size, gap_open=-10, gap_extend=-1, seed=1234):
    """C4 profiles for the library side: consensus = the reference side of each pair, PSSM = BLOSUM62 row + noise,
    gap open/close set for positions 1..=len like examples/pssm_bench.rs:67-83 (position 0 keeps -128)."""
    b62 = lib.builtin_matrix("BLOSUM62")[1]
    rng = np.random.default_rng(seed)
    out = []
    n = len(ro) - 1
    for k in range(n):
        cons = ra[int(ro[k]):int(ro[k + 1])].tobytes()
        p = api.AAProfile(lib, len(cons), block_size, gap_extend)
        if len(cons):
            p.set_all(MAP20, pssm_scores(b62, cons, rng))
        p.set_all_gap_open_C(gap_open)
        p.set_all_gap_close_C(0)
        p.set_all_gap_open_R(gap_open)
        p.set_gap_open_C(0, -128)
        p.set_gap_close_C(0, -128)
        p.set_gap_open_R(0, -128)
        out.append(p)
    return out


def make_pssm_batch(lib, ra, ro, gap_open=-10, gap_extend=-1, seed=1234, pinned=False):
    """The C4 profiles of make_lib_profiles as one api.PssmBatch (built on the device): same numbers, same rng order.
    Per-position gap arrays with -128 at position 0 (the reference loader skips i == 0, examples/pssm_bench.rs:67-83)."""
    b62 = lib.builtin_matrix("BLOSUM62")[1]
    rng = np.random.default_rng(seed)
    n = len(ro) - 1
    lens = (ro[1:] - ro[:-1]).astype(np.uint64)
    soff = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens * np.uint64(20), out=soff[1:])
    goff = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens + np.uint64(1), out=goff[1:])

    def buf(nbytes):
        if pinned:
            import torch
            t = torch.empty(max(int(nbytes), 1), dtype=torch.int8).pin_memory()
            return t.numpy()[:int(nbytes)], t
        return np.zeros(int(nbytes), dtype=np.int8), None

    scores, k0 = buf(soff[-1])
    go, k1 = buf(goff[-1]); gc, k2 = buf(goff[-1]); gr, k3 = buf(goff[-1])
    for k in range(n):
        cons = ra[int(ro[k]):int(ro[k + 1])].tobytes()
        if len(cons):
            scores[int(soff[k]):int(soff[k + 1])] = pssm_scores(b62, cons, rng).reshape(-1)
        a, e = int(goff[k]), int(goff[k + 1])
        go[a:e] = gap_open; gc[a:e] = 0; gr[a:e] = gap_open
        go[a] = gc[a] = gr[a] = -128
    pb = api.PssmBatch(MAP20, scores, soff, gap_extend, gaps=(go, gc, gr))
    pb._pinned = (k0, k1, k2, k3)
    return pb
