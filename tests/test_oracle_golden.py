"""Pins the CPU oracle against every known-answer assertion the reference's own tests hold for the
hot path: /root/reference/src/scan_block.rs:1908-2230 (58 assertions), src/avx2.rs:470-488 (prefix
scan vectors) and the doc-test src/lib.rs:8-35.  Inputs are copied as data (byte strings and
expected numbers); nothing from /root/reference is read at run time.
"""
import numpy as np
import pytest

import ora
from ora import AA, BYTE, NUC, TRACE, XDROP, LOCAL_START, FREE_QUERY_START_GAPS, FREE_QUERY_END_GAPS
from ora import Block, Padded, Profile


def B62():
    return ora.builtin("BLOSUM62")


# ---- scan_block.rs:1908-1992 test_no_x_drop ------------------------------------------------------
NO_X_DROP_AA = [  # (q, r, score) with BLOSUM62, gaps -11/-1, 16..=16
    (b"", b"", 0), (b"", b"AAAA", -14), (b"AAAA", b"", -14), (b"AARA", b"AAAA", 11),
    (b"AARAAAA", b"AAAAAAAA", 12), (b"AAAA", b"AAAA", 16), (b"AARA", b"AAAA", 11),
    (b"RRRR", b"AAAA", -4), (b"AAA", b"AAAA", 1),
]
NO_X_DROP_NUC = [  # NW1, gaps -2/-1, 16..=16
    (b"ATAA", b"AAAN", 0), (b"A" * 32, b"A" * 32, 32), (b"T" * 32, b"A" * 32, -32),
    (b"TA" * 16, b"A" * 32, 0), (b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", 7),
    (b"C", b"AAAA", -5), (b"AAAA", b"C", -5),
]


def test_no_x_drop():
    a = Block(100, 100, 16, 0)
    for q, r, s in NO_X_DROP_AA:
        assert a.align(Padded(AA, q, 16), Padded(AA, r, 16), AA, B62(), (-11, -1), (16, 16), 0)[0] == s, (q, r)
    for q, r, s in NO_X_DROP_NUC:
        assert a.align(Padded(NUC, q, 16), Padded(NUC, r, 16), NUC, ora.nw1(), (-2, -1), (16, 16), 0)[0] == s, (q, r)


# ---- scan_block.rs:1994-2050 test_x_drop ---------------------------------------------------------
def test_x_drop():
    g = (-11, -1)
    a = Block(100, 100, 16, XDROP)
    for q, r, exp in [(b"", b"", (0, 0, 0)), (b"", b"AAAA", (0, 0, 0)), (b"AAAA", b"", (0, 0, 0)),
                      (b"AAAAAA", b"AAARRA", (14, 6, 6)),
                      (b"A" * 44, b"A" * 15 + b"R" * 16 + b"A" * 13, (60, 15, 15))]:
        assert a.align(Padded(AA, q, 16), Padded(AA, r, 16), AA, B62(), g, (16, 16), 1) == exp
    a = Block(2048, 2048, 2048, TRACE | XDROP)
    s = b"A" * 2048
    assert a.align(Padded(AA, s, 2048), Padded(AA, s, 2048), AA, B62(), g, (2048, 2048), 100) == (8192, 2048, 2048)
    a = Block(0, 0, 16, TRACE | XDROP)
    assert a.align(Padded(AA, b"", 16), Padded(AA, b"", 16), AA, B62(), g, (16, 16), 1) == (0, 0, 0)
    a = Block(4, 4, 16, TRACE | XDROP)
    assert a.align(Padded(AA, b"", 16), Padded(AA, b"AAAA", 16), AA, B62(), g, (16, 16), 1) == (0, 0, 0)
    assert a.align(Padded(AA, b"AAAA", 16), Padded(AA, b"", 16), AA, B62(), g, (16, 16), 1) == (0, 0, 0)


# ---- scan_block.rs:2052-2103 test_trace ----------------------------------------------------------
def test_trace():
    a = Block(100, 100, 16, TRACE)
    res = a.align(Padded(AA, b"AAAAAA", 16), Padded(AA, b"AAARRA", 16), AA, B62(), (-11, -1), (16, 16), 0)
    assert res == (14, 6, 6)
    assert a.cigar(6, 6, eq=True) == "3=2X1="
    res = a.align(Padded(AA, b"AAA", 16), Padded(AA, b"AAAA", 16), AA, B62(), (-11, -1), (16, 16), 0)
    assert res == (1, 3, 4)
    assert a.cigar(3, 4) == "3M1D"
    res = a.align(Padded(NUC, b"TTTTTTTTAAAAAAATTTTTTTTT", 16), Padded(NUC, b"TTAAAAAAATTTTTTTTTTTT", 16), NUC, ora.nw1(),
                  (-2, -1), (16, 16), 0)
    assert res == (7, 24, 21)
    assert a.cigar(24, 21) == "2M6I16M3D"
    a = Block(100, 100, 32, TRACE)
    q, r = Padded(NUC, b"AAAAAAAAATTGCGCT", 32), Padded(NUC, b"AAAAAAAAAGCGC", 32)
    assert a.align(q, r, NUC, ora.nw1(), (-2, -1), (32, 32), 0) == (8, 16, 13)
    assert a.cigar(16, 13, eq=True) == "9=2I4=1I"
    assert a.align(q, r, NUC, ora.nuc_matrix(2, -1), (-5, -2), (32, 32), 0) == (14, 16, 13)
    assert a.cigar(16, 13, eq=True) == "9=2I4=1I"


# ---- scan_block.rs:2105-2120 test_bytes ----------------------------------------------------------
def test_bytes():
    a = Block(100, 100, 16, 0)
    m = np.array([1, -1], dtype=np.int8)  # BYTES1
    assert a.align(Padded(BYTE, b"AAAAAA", 16), Padded(BYTE, b"AAAaaA", 16), BYTE, m, (-2, -1), (16, 16), 0)[0] == 2
    assert a.align(Padded(BYTE, b"abdefg", 16), Padded(BYTE, b"abcdefg", 16), BYTE, m, (-2, -1), (16, 16), 0)[0] == 4


# ---- scan_block.rs:2122-2168 test_profile --------------------------------------------------------
def test_profile():
    a = Block(100, 100, 16, 0)
    q = Padded(AA, b"AAAA", 16)
    assert a.align_profile(q, Profile.from_bytes(b"AAAA", 16, 1, -1, -1, 0, -1, -1), (16, 16), 0)[0] == 4
    assert a.align_profile(q, Profile.from_bytes(b"AATTAA", 16, 1, -1, -1, 0, -1, -1), (16, 16), 0)[0] == 1
    assert a.align_profile(q, Profile.from_bytes(b"AATTAA", 16, 1, -1, -1, -1, -1, -1), (16, 16), 0)[0] == 0
    a = Block(100, 100, 16, TRACE)
    q = Padded(AA, b"TTTTTTTTAAAAAAATTTTTTTTT", 16)
    r = Profile.from_bytes(b"TTAAAAAAATTTTTTTTTTTT", 16, 1, -1, -1, 0, -1, -1)
    assert a.align_profile(q, r, (16, 16), 0) == (7, 24, 21)
    assert a.cigar(24, 21) == "2M6I16M3D"
    r = Profile.from_bytes(b"TTAAAAAAATTTTTTTTTTTT", 16, 1, -1, -1, -1, -1, -1)
    assert a.align_profile(q, r, (16, 16), 0) == (6, 24, 21)
    assert a.cigar(24, 21) == "2M6I16M3D"
    r = Profile.from_bytes(b"TTAAAAAAATTTTTTTTTTTT", 16, 1, -1, -2, -1, -1, -1)
    r.set_gap_close_C(17, -1)
    r.set_gap_close_C(19, 0)
    assert a.align_profile(q, r, (16, 16), 0) == (6, 24, 21)
    assert a.cigar(24, 21) == "2M6I14M3D2M"


# ---- scan_block.rs:2170-2230 test_local_and_free_query_gaps --------------------------------------
def test_local_and_free_query_gaps():
    g = (-2, -1)
    local = Block(100, 100, 32, TRACE | LOCAL_START)
    q, r = Padded(NUC, b"CCCCCCCCCCAAAAAA", 32), Padded(NUC, b"TTTTAAAAAA", 32)
    assert local.align(q, r, NUC, ora.nw1(), g, (32, 32), 0) == (6, 16, 10)
    assert local.cigar(16, 10, eq=True) == "6="
    local = Block(100, 100, 32, TRACE | XDROP | LOCAL_START)
    q, r = Padded(NUC, b"CCCCCCCCCCAAAAAACCCCCCCCCCCC", 32), Padded(NUC, b"TTTTAAAAAATTTTTTT", 32)
    assert local.align(q, r, NUC, ora.nw1(), g, (32, 32), 100) == (6, 16, 10)
    assert local.cigar(16, 10, eq=True) == "6="
    qs = Block(100, 100, 32, TRACE | FREE_QUERY_START_GAPS)
    q, r = Padded(NUC, b"AAAAAA", 32), Padded(NUC, b"CCCCCCCCCCAAAAAA", 32)
    assert qs.align(q, r, NUC, ora.nw1(), g, (32, 32), 0) == (6, 6, 16)
    assert qs.cigar(6, 16, eq=True) == "6="
    r = Padded(NUC, b"CCCCCCCCCCAAATAA", 32)
    assert qs.align(q, r, NUC, ora.nw1(), g, (32, 32), 0) == (4, 6, 16)
    assert qs.cigar(6, 16, eq=True) == "3=1X2="
    qe = Block(100, 100, 32, TRACE | FREE_QUERY_END_GAPS)
    r = Padded(NUC, b"AAAAAACCCCCCCCCC", 32)
    assert qe.align(q, r, NUC, ora.nw1(), g, (32, 32), 0) == (6, 6, 6)
    assert qe.cigar(6, 6, eq=True) == "6="
    r = Padded(NUC, b"AAATAACCCCCCCCCC", 32)
    assert qe.align(q, r, NUC, ora.nw1(), g, (32, 32), 0) == (4, 6, 6)
    assert qe.cigar(6, 6, eq=True) == "3=1X2="


# ---- lib.rs:8-35 doc-test ------------------------------------------------------------------------
def test_doc_example():
    q = Padded(NUC, b"TTTTTTTTAAAAAAATTTTTTTTT", 256)
    r = Padded(NUC, b"TTAAAAAAATTTTTTTTTTTT", 256)
    a = Block(len(q), len(r), 256, TRACE)
    assert a.align(q, r, NUC, ora.nw1(), (-2, -1), (32, 256), 0) == (7, 24, 21)
    assert a.cigar(24, 21, eq=True) == "2=6I16=3D"


# ---- avx2.rs:470-488 test_prefix_scan ------------------------------------------------------------
def test_prefix_scan_vectors():
    v = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 12, 13, 14, 11]
    assert list(ora.prefix_scan(v, 0)) == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 15, 15, 15, 15]
    assert list(ora.prefix_scan(v, -1)) == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 14, 13, 14, 13]


# ---- c/example.c (the three examples print these) ------------------------------------------------
def test_c_example_values():
    a = Block(8, 7, 32, TRACE)
    q, r = Padded(AA, b"AAAAAAAA", 32), Padded(AA, b"AARAAAA", 32)
    res = a.align(q, r, AA, B62(), (-11, -1), (32, 32), 0)
    # full-matrix affine NW: 7 matches minus one R/A mismatch, one gap of length 1
    assert res[1:] == (8, 7)
    # re-score the CIGAR to make sure it is consistent with the score
    runs = a.cigar_runs(8, 7, eq=False)
    assert sum(n for op, n in runs if op in (1, 4)) == 8 and sum(n for op, n in runs if op in (1, 5)) == 7


def test_argument_rules():
    # scan_block.rs:849-862
    a = Block(10, 10, 32, 0)
    q = Padded(NUC, b"ACGT", 32)
    with pytest.raises(ValueError):
        a.align(q, q, NUC, ora.nw1(), (1, -1), (32, 32), 0)      # gaps must be negative
    with pytest.raises(ValueError):
        a.align(q, q, NUC, ora.nw1(), (-1, -1), (32, 32), 0)     # open < extend
    with pytest.raises(ValueError):
        a.align(q, q, NUC, ora.nw1(), (-2, -1), (24, 32), 0)     # powers of two
    with pytest.raises(ValueError):
        a.align(q, q, NUC, ora.nw1(), (-2, -1), (32, 64), 0)     # larger than the Block was made for
    with pytest.raises(ValueError):
        Block(10, 10, 48, 0)                                       # scan_block.rs:799
    # sizes below L are clamped up to 16 (scan_block.rs:853-854)
    assert a.align(q, q, NUC, ora.nw1(), (-2, -1), (4, 8), 0) == (4, 4, 4)


def test_golden_fixture_file_matches_and_oracle_reproduces_it():
    """tests/golden/reference_unit_tests.json (made by tests/golden/make_golden.py) is in step with golden_cases.py,
    and the oracle reproduces every vector in it when driven from the file alone."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_unit_tests.json")
    doc = json.load(open(path))
    import golden_cases as G
    assert len(doc["cases"]) == len(G.SEQ_CASES)
    kinds = {"AAMatrix": ora.AA, "NucMatrix": ora.NUC, "ByteMatrix": ora.BYTE}
    for c in doc["cases"]:
        kind = kinds[c["matrix_kind"]]
        if c["matrix"] == "NW1":
            m = ora.nw1()
        elif c["matrix"] == "BYTES1":
            m = np.array([1, -1], dtype=np.int8)
        elif isinstance(c["matrix"], str):
            m = ora.builtin(c["matrix"])
        else:
            m = ora.nuc_matrix(*c["matrix"])
        pad = c["size"][1]
        q, r = ora.Padded(kind, c["query"].encode(), pad), ora.Padded(kind, c["reference"].encode(), pad)
        b = ora.Block(len(q), len(r), pad, (ora.TRACE if c["trace"] else 0) | (ora.XDROP if c["x_drop_mode"] else 0))
        res = b.align(q, r, kind, m, tuple(c["gaps"]), tuple(c["size"]), c["x_drop"])
        assert res == (c["score"], c["query_idx"], c["reference_idx"]), c
        if c["cigar"] is not None:
            assert b.cigar(res[1], res[2], c["cigar_eq"]) == c["cigar"], c
