"""Host-side logic of the C ABI, exercised through the emulated build (no GPU): argument rules of
Block::align (scan_block.rs:849-862), alphabet checks (scores.rs:130-134, 212-216), helper functions."""
import ctypes as C

import numpy as np
import pytest

import backend
from block_aligner_b200 import api


@pytest.fixture(scope="module")
def env():
    lib = backend.emu_lib()
    return lib, api.Aligner(lib)


def _run(al, q, r, gaps=(-2, -1), size=(32, 32), x_drop=0, flags=0, scoring=api.SCORING_NUC, matrix=None):
    matrix = api.nuc_matrix(1, -1) if matrix is None else matrix
    return al.align_batch([q], [r], scoring, matrix, gaps, size, x_drop, flags)[0][0]


def test_argument_rules(env):
    lib, al = env
    with pytest.raises(api.BlockAlignerError, match="negative"):
        _run(al, b"ACGT", b"ACGT", gaps=(1, -1))
    with pytest.raises(api.BlockAlignerError, match="more than gap extend"):
        _run(al, b"ACGT", b"ACGT", gaps=(-1, -1))
    with pytest.raises(api.BlockAlignerError, match="powers of two"):
        _run(al, b"ACGT", b"ACGT", size=(24, 32))
    with pytest.raises(api.BlockAlignerError, match="nonnegative"):
        _run(al, b"ACGT", b"ACGT", x_drop=-1, flags=api.XDROP)
    with pytest.raises(api.BlockAlignerError, match="16384"):
        _run(al, b"ACGT", b"ACGT", size=(32, 32768))
    with pytest.raises(api.BlockAlignerError, match="larger than max"):      # our rule: the reference walks off its scratch here
        _run(al, b"ACGT", b"ACGT", size=(64, 32))
    assert _run(al, b"ACGT", b"ACGT", size=(4, 8)) == (4, 4, 4)     # sizes below L are clamped up to 16


def test_alphabet_checks(env):
    lib, al = env
    assert _run(al, b"acgt", b"ACGT") == (4, 4, 4)                   # lowercase is uppercased
    with pytest.raises(api.BlockAlignerError, match="alphabet"):
        _run(al, b"AC-T", b"ACGT")
    b62 = lib.builtin_matrix("BLOSUM62")[1]
    with pytest.raises(api.BlockAlignerError, match="alphabet"):
        _run(al, b"AC*T", b"ACGT", scoring=api.SCORING_AA, matrix=b62, gaps=(-11, -1))


def test_empty_batch(env):
    lib, al = env
    res, cig, st = al.align_batch([], [], api.SCORING_NUC, api.nuc_matrix(1, -1), (-2, -1), (32, 32))
    assert res == [] and st.cells == 0


def test_percent_len(env):
    lib, _ = env
    # lib.rs:109-111: round(p * len) -> max 32 -> next power of two -> min 16384
    assert lib.percent_len(10_000, 0.01) == 128
    assert lib.percent_len(50_000, 0.1) == 8192
    assert lib.percent_len(100, 0.01) == 32
    assert lib.percent_len(10_000_000, 0.5) == 16384


def test_cigar_format(env):
    lib, _ = env
    runs = np.array([(9 << 4) | 2, (2 << 4) | 4, (4 << 4) | 2, (1 << 4) | 4], dtype=np.uint32)
    buf = C.create_string_buffer(64)
    n = lib.L.ba_cigar_format(runs.ctypes.data, len(runs), buf, 64)
    assert buf.value == b"9=2I4=1I" and n == 8


def test_builtin_matrices_match_reference_layout(env):
    lib, _ = env
    # NW1 = NucMatrix::new_simple(1, -1) (scores.rs:277): index (a & 7) * 16 + (b & 15)
    assert (lib.builtin_matrix("NW1")[1] == api.nuc_matrix(1, -1)).all()
    b62 = lib.builtin_matrix("BLOSUM62")[1].reshape(27, 32)
    assert b62[0, 0] == 4 and b62[ord("W") - 65, ord("W") - 65] == 11 and b62[ord("A") - 65, ord("R") - 65] == -1
    assert (b62[26] == -128).all() and (b62[:, 26:] == -128).all()
    assert (b62[:26, :26] == b62[:26, :26].T).all()
    for name in ("BLOSUM45", "BLOSUM50", "BLOSUM80", "BLOSUM90", "PAM100", "PAM120", "PAM160", "PAM200", "PAM250"):
        m = lib.builtin_matrix(name)[1].reshape(27, 32)
        assert (m[:26, :26] == m[:26, :26].T).all() and (m[26] == -128).all()


def test_align_batch_pipelined_chunks(env, monkeypatch):
    """ba_align_batch cuts big batches into pipelined chunks; results must not depend on the cut."""
    import parity
    from block_aligner_b200 import workloads
    lib, al = env
    w = workloads.WORKLOADS["C2_nanopore_xdrop_10k"]
    gen = workloads.params(alphabet=0, len_dist=0, len_min=100, len_max=900, sub_rate=0.04, ins_rate=0.04, del_rate=0.04, suffix_len=60)
    qa, qo, ra, ro = workloads.generate(gen, 37, stream=2)
    m = workloads.matrix_of(lib, w)
    ref = parity.run_lib(lib, al, w["scoring"], m, w["gaps"], w["size"], w["x_drop"], w["flags"], False, qa, qo, ra, ro)
    cfg = al.config(w["scoring"], m, w["gaps"], w["size"], w["x_drop"], w["flags"], False)
    for chunks, geom in (("1", "0"), ("3", "0"), ("64", "0"), ("5", "1"), ("12", "1")):
        monkeypatch.setenv("BA_PIPELINE_CHUNKS", chunks)
        monkeypatch.setenv("BA_PIPELINE_GEOM", geom)     # geometric chunk sizes (the default for large batches)
        out = np.zeros(37, dtype=np.dtype([("score", np.int32), ("q", np.uint64), ("r", np.uint64)], align=True))
        st = api.BaStats()
        lib.check(lib.L.ba_align_batch(al.h, C.byref(cfg), 37, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data,
                                       out.ctypes.data, C.byref(st)))
        assert (out["score"] == ref[0][:, 0]).all() and (out["q"] == ref[0][:, 1]).all() and (out["r"] == ref[0][:, 2]).all()
        assert st.cells == int(ref[1].sum()) and st.n_failed == 0


def test_align_exp_matches_reference_loop(env):
    """ba_align_batch_exp vs the reference's align_exp loop (scan_block.rs:884-902) run on the oracle."""
    import ora
    from block_aligner_b200 import workloads
    lib, al = env
    gen = workloads.params(alphabet=0, len_dist=0, len_min=300, len_max=1200, sub_rate=0.05, ins_rate=0.03, del_rate=0.03,
                           big_indel_prob=0.9, big_indel_min=40, big_indel_max=150)
    qa, qo, ra, ro = workloads.generate(gen, 16, stream=51)
    qs = [qa[int(qo[k]):int(qo[k + 1])].tobytes() for k in range(16)]
    rs = [ra[int(ro[k]):int(ro[k + 1])].tobytes() for k in range(16)]
    nw1 = lib.builtin_matrix("NW1")[1]
    # targets: the score a 256-wide block finds (so small blocks often miss it)
    full, _, _ = al.align_batch(qs, rs, api.SCORING_NUC, nw1, (-2, -1), (256, 256))
    targets = [f[0] for f in full]
    res, used, _ = api.align_batch_exp(al, qs, rs, api.SCORING_NUC, nw1, (-2, -1), (32, 256), targets)
    for k in range(16):
        ob = ora.Block(len(qs[k]), len(rs[k]), 256, 0)
        q, r = ora.Padded(ora.NUC, qs[k], 256), ora.Padded(ora.NUC, rs[k], 256)
        exp_used, exp_res, mn = None, None, 32
        while mn <= 256:
            exp_res = ob.align(q, r, ora.NUC, ora.nw1(), (-2, -1), (mn, 256), 0)
            if exp_res[0] >= targets[k]:
                exp_used = mn
                break
            mn *= 2
        assert res[k] == exp_res and used[k] == exp_used, (k, res[k], exp_res, used[k], exp_used)
    assert any(u != 32 for u in used), "test inputs should need at least one retry"


def test_align_batch_cigar_delivers_runs_into_caller_buffer(env, monkeypatch):
    """ba_align_batch_cigar == upload/run/download + ba_batch_cigar for every pair, across pipelined chunks"""
    import parity
    from block_aligner_b200 import workloads
    lib, al = env
    gen = workloads.params(alphabet=0, len_dist=0, len_min=100, len_max=900, sub_rate=0.04, ins_rate=0.04, del_rate=0.04, suffix_len=60)
    n = 41
    qa, qo, ra, ro = workloads.generate(gen, n, stream=2)
    m = lib.builtin_matrix("NW1")[1]
    flags = api.TRACE | api.XDROP
    ref = parity.run_lib(lib, al, api.SCORING_NUC, m, (-2, -1), (32, 256), 50, flags, True, qa, qo, ra, ro)
    cfg = al.config(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, flags, True)
    for chunks in ("1", "4"):
        monkeypatch.setenv("BA_PIPELINE_CHUNKS", chunks)
        out = np.zeros(n, dtype=np.dtype([("score", np.int32), ("q", np.uint64), ("r", np.uint64)], align=True))
        cap = int(qo[-1] + ro[-1]) + 5 * n
        runs = np.zeros(cap, dtype=np.uint32)
        off = np.zeros(n, dtype=np.uint64)
        ln = np.zeros(n, dtype=np.uint32)
        used = C.c_size_t()
        st = api.BaStats()
        lib.check(lib.L.ba_align_batch_cigar(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data,
                                             out.ctypes.data, runs.ctypes.data, cap, off.ctypes.data, ln.ctypes.data,
                                             C.byref(used), C.byref(st)))
        assert (out["score"] == ref[0][:, 0]).all()
        assert used.value == int(ln.sum())
        for k in range(n):
            got = runs[int(off[k]):int(off[k]) + int(ln[k])].astype(np.uint64)
            assert len(got) == len(ref[2][k]) and (got == ref[2][k]).all()
    # a buffer that is too small is an error, not a truncation
    small = np.zeros(8, dtype=np.uint32)
    rc = lib.L.ba_align_batch_cigar(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data,
                                    out.ctypes.data, small.ctypes.data, 8, off.ctypes.data, ln.ctypes.data, C.byref(used), C.byref(st))
    assert rc == api.ERR_OVERFLOW if hasattr(api, "ERR_OVERFLOW") else rc != 0


def test_align_profile_exp_matches_reference_loop(env):
    """ba_align_batch_exp_pssm / _profiles vs the reference's align_profile_exp loop (scan_block.rs:974-992) on the oracle"""
    import ora
    import parity
    from block_aligner_b200 import workloads
    lib, al = env
    w = workloads.WORKLOADS["C4_seq_to_profile_xdrop"]
    n = 20
    gen = workloads.params(alphabet=1, len_dist=0, len_min=300, len_max=700, sub_rate=0.3, ins_rate=0.03, del_rate=0.03,
                           big_indel_prob=1.0, big_indel_min=60, big_indel_max=120)
    qa, qo, ra, ro = workloads.generate(gen, n, seed=3, stream=4)
    qs = [qa[int(qo[k]):int(qo[k + 1])].tobytes() for k in range(n)]
    pb = workloads.make_pssm_batch(lib, ra, ro, seed=21)
    oprofs = parity.make_ora_profiles(ra, ro, 256, -10, -1, 21)
    # targets: what a 256-wide block finds
    exp_full = parity.oracle_batch(api.SCORING_PROFILE, None, None, (256, 256), 0, 0, False, qa, qo, ra, ro, profiles=oprofs)
    targets = [int(x) for x in exp_full[0][:, 0]]
    targets[3] = 10 ** 6
    res, used, st = api.align_batch_exp(al, qs, None, api.SCORING_PROFILE, None, None, (32, 256), targets, profiles=pb)
    lprofs = workloads.make_lib_profiles(lib, ra, ro, 256, seed=21)
    res2, used2, _ = api.align_batch_exp(al, qs, None, api.SCORING_PROFILE, None, None, (32, 256), targets, profiles=lprofs)
    assert res == res2 and used == used2
    cells = 0
    for k in range(n):
        ob = ora.Block(len(qs[k]), len(oprofs[k]), 256, 0)
        q = ora.Padded(ora.AA, qs[k], 256)
        exp_used, exp_res, mn = None, None, 32
        while mn <= 256:
            exp_res = ob.align_profile(q, oprofs[k], (mn, 256), 0)
            cells += ob.cells()
            if exp_res[0] >= targets[k]:
                exp_used = mn
                break
            mn *= 2
        assert res[k] == exp_res and used[k] == exp_used, (k, res[k], exp_res, used[k], exp_used)
    assert used[3] is None and st.cells == cells and st.n_failed == 0
    assert any(u not in (32, None) for u in used), "test inputs should need at least one retry"


def test_align_exp_rejects_min_above_max(env):
    lib, al = env
    with pytest.raises(api.BlockAlignerError, match="larger than max"):
        api.align_batch_exp(al, [b"ACGT"], [b"ACGT"], api.SCORING_NUC, api.nuc_matrix(1, -1), (-2, -1), (64, 32), [1])


def test_align_batch_pssm_pipelined_chunks(env, monkeypatch):
    """ba_align_batch_pssm cuts big batches into pipelined chunks; results must not depend on the cut."""
    import parity
    from block_aligner_b200 import workloads
    lib, al = env
    w = workloads.WORKLOADS["C4_seq_to_profile_xdrop"]
    n = 23
    qa, qo, ra, ro = workloads.generate(w["gen"], n, seed=8, stream=4)
    cfg = al.config(api.SCORING_PROFILE, None, None, (32, 128), w["x_drop"], w["flags"], False)
    exp = parity.oracle_batch(api.SCORING_PROFILE, None, None, (32, 128), w["x_drop"], w["flags"], False, qa, qo, ra, ro,
                              profiles=parity.make_ora_profiles(ra, ro, 128, -10, -1, 77))
    pb = workloads.make_pssm_batch(lib, ra, ro, seed=77)
    for chunks in ("1", "2", "5"):
        monkeypatch.setenv("BA_PIPELINE_CHUNKS", chunks)
        out = np.zeros(n, dtype=parity.ABI_RES_DT)
        st = api.BaStats()
        lib.check(lib.L.ba_align_batch_pssm(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, C.byref(pb.c), out.ctypes.data, C.byref(st)))
        res = np.stack([out["score"].astype(np.int64), out["q"].astype(np.int64), out["r"].astype(np.int64)], axis=1)
        assert parity.compare_abi(f"pssm-chunks{chunks}", res, None, st, exp) == 0
