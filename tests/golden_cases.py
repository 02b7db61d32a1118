"""Known-answer cases of the reference (data only: byte strings and expected numbers), shared by the
emulated (CPU) and CUDA (GPU) test modules. Sources: /root/reference/src/scan_block.rs:1908-2168,
src/lib.rs:8-35."""
import ctypes as C

import ora
from block_aligner_b200 import api

AA, NUC, BYTE = api.SCORING_AA, api.SCORING_NUC, api.SCORING_BYTE
T, X = api.TRACE, api.XDROP

# (scoring, matrix, gaps, size, x_drop, flags, cigar_eq, q, r, expected result, expected cigar or None)
SEQ_CASES = []
for q, r, s in [(b"", b"", 0), (b"", b"AAAA", -14), (b"AAAA", b"", -14), (b"AARA", b"AAAA", 11),
                (b"AARAAAA", b"AAAAAAAA", 12), (b"AAAA", b"AAAA", 16), (b"RRRR", b"AAAA", -4), (b"AAA", b"AAAA", 1)]:
    SEQ_CASES.append((AA, "BLOSUM62", (-11, -1), (16, 16), 0, 0, False, q, r, (s, len(q), len(r)), None))
for q, r, s in [(b"ATAA", b"AAAN", 0), (b"A" * 32, b"A" * 32, 32), (b"T" * 32, b"A" * 32, -32), (b"TA" * 16, b"A" * 32, 0),
                (b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", 7), (b"C", b"AAAA", -5), (b"AAAA", b"C", -5)]:
    SEQ_CASES.append((NUC, "NW1", (-2, -1), (16, 16), 0, 0, False, q, r, (s, len(q), len(r)), None))
for q, r, e in [(b"", b"", (0, 0, 0)), (b"", b"AAAA", (0, 0, 0)), (b"AAAA", b"", (0, 0, 0)), (b"AAAAAA", b"AAARRA", (14, 6, 6)),
                (b"A" * 44, b"A" * 15 + b"R" * 16 + b"A" * 13, (60, 15, 15))]:
    SEQ_CASES.append((AA, "BLOSUM62", (-11, -1), (16, 16), 1, X, False, q, r, e, None))
SEQ_CASES += [
    (AA, "BLOSUM62", (-11, -1), (2048, 2048), 100, T | X, False, b"A" * 2048, b"A" * 2048, (8192, 2048, 2048), None),
    (AA, "BLOSUM62", (-11, -1), (16, 16), 0, T, True, b"AAAAAA", b"AAARRA", (14, 6, 6), "3=2X1="),
    (AA, "BLOSUM62", (-11, -1), (16, 16), 0, T, False, b"AAA", b"AAAA", (1, 3, 4), "3M1D"),
    (NUC, "NW1", (-2, -1), (16, 16), 0, T, False, b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", (7, 24, 21), "2M6I16M3D"),
    (NUC, "NW1", (-2, -1), (32, 32), 0, T, True, b"AAAAAAAAATTGCGCT", b"AAAAAAAAAGCGC", (8, 16, 13), "9=2I4=1I"),
    (NUC, (2, -1), (-5, -2), (32, 32), 0, T, True, b"AAAAAAAAATTGCGCT", b"AAAAAAAAAGCGC", (14, 16, 13), "9=2I4=1I"),
    (BYTE, "BYTES1", (-2, -1), (16, 16), 0, 0, False, b"AAAAAA", b"AAAaaA", (2, 6, 6), None),
    (BYTE, "BYTES1", (-2, -1), (16, 16), 0, 0, False, b"abdefg", b"abcdefg", (4, 6, 7), None),
    # doc-test lib.rs:8-35
    (NUC, "NW1", (-2, -1), (32, 256), 0, T, True, b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", (7, 24, 21), "2=6I16=3D"),
]


def _matrix(lib, m, scoring):
    if isinstance(m, str):
        return lib.builtin_matrix(m)[1]
    return api.nuc_matrix(*m) if scoring == NUC else api.aa_matrix_simple(*m)


def run_batch_golden(lib, al):
    for (sc, m, gaps, size, xd, flags, eq, q, r, exp, cig) in SEQ_CASES:
        res, cigs, _ = al.align_batch([q], [r], sc, _matrix(lib, m, sc), gaps, size, xd, flags, eq)
        assert res[0] == exp, (q[:20], r[:20], res[0], exp)
        if cig is not None:
            assert cigs[0] == cig, (q, r, cigs[0], cig)
    # the same cases as ONE batch per configuration (exercises the ticket queue / several warps)
    groups = {}
    for c in SEQ_CASES:
        groups.setdefault((c[0], str(c[1]), c[2], c[3], c[4], c[5], c[6]), []).append(c)
    for key, cs in groups.items():
        sc, _, gaps, size, xd, flags, eq = key
        res, cigs, _ = al.align_batch([c[7] for c in cs], [c[8] for c in cs], sc, _matrix(lib, cs[0][1], sc), gaps, size, xd, flags, eq)
        for k, c in enumerate(cs):
            assert res[k] == c[9]
            if c[10] is not None:
                assert cigs[k] == c[10]


def run_legacy_golden(lib):
    """Part 1 of the C ABI (what c/example.c uses): scan_block.rs test_no_x_drop / test_x_drop / test_trace /
    test_profile, amino-acid monomorphs only (the reference's C API has no nucleotide bindings)."""
    b62 = C.addressof((C.c_int8 * 864).in_dll(lib.L, "BLOSUM62"))
    g = (-11, -1)
    a = api.Block(lib, 100, 100, 16)
    for q, r, s in [(b"", b"", 0), (b"", b"AAAA", -14), (b"AAAA", b"", -14), (b"AARA", b"AAAA", 11),
                    (b"AARAAAA", b"AAAAAAAA", 12), (b"AAAA", b"AAAA", 16), (b"RRRR", b"AAAA", -4), (b"AAA", b"AAAA", 1)]:
        assert a.align(api.PaddedBytes(lib, q, 16), api.PaddedBytes(lib, r, 16), b62, g, (16, 16), 0)[0] == s
    a = api.Block(lib, 100, 100, 16, x_drop=True)
    assert a.align(api.PaddedBytes(lib, b"AAAAAA", 16), api.PaddedBytes(lib, b"AAARRA", 16), b62, g, (16, 16), 1) == (14, 6, 6)
    assert a.align(api.PaddedBytes(lib, b"A" * 44, 16), api.PaddedBytes(lib, b"A" * 15 + b"R" * 16 + b"A" * 13, 16), b62, g,
                   (16, 16), 1) == (60, 15, 15)
    a = api.Block(lib, 100, 100, 16, trace=True)
    cg = api.Cigar(lib, 100, 100)
    q, r = api.PaddedBytes(lib, b"AAAAAA", 16), api.PaddedBytes(lib, b"AAARRA", 16)
    assert a.align(q, r, b62, g, (16, 16), 0) == (14, 6, 6)
    assert a.cigar_eq(q, r, 6, 6, cg) == "3=2X1="
    q, r = api.PaddedBytes(lib, b"AAA", 16), api.PaddedBytes(lib, b"AAAA", 16)
    assert a.align(q, r, b62, g, (16, 16), 0) == (1, 3, 4)
    assert a.cigar(3, 4, cg) == "3M1D"
    # test_profile
    a0 = api.Block(lib, 100, 100, 16)
    q = api.PaddedBytes(lib, b"AAAA", 16)
    assert a0.align_profile(q, api.AAProfile.from_bytes(lib, b"AAAA", 16, 1, -1, -1, 0, -1, -1), (16, 16), 0)[0] == 4
    assert a0.align_profile(q, api.AAProfile.from_bytes(lib, b"AATTAA", 16, 1, -1, -1, 0, -1, -1), (16, 16), 0)[0] == 1
    assert a0.align_profile(q, api.AAProfile.from_bytes(lib, b"AATTAA", 16, 1, -1, -1, -1, -1, -1), (16, 16), 0)[0] == 0
    q = api.PaddedBytes(lib, b"TTTTTTTTAAAAAAATTTTTTTTT", 16)
    p = api.AAProfile.from_bytes(lib, b"TTAAAAAAATTTTTTTTTTTT", 16, 1, -1, -1, 0, -1, -1)
    assert a.align_profile(q, p, (16, 16), 0) == (7, 24, 21)
    assert a.cigar(24, 21, cg) == "2M6I16M3D"
    p = api.AAProfile.from_bytes(lib, b"TTAAAAAAATTTTTTTTTTTT", 16, 1, -1, -1, -1, -1, -1)
    assert a.align_profile(q, p, (16, 16), 0) == (6, 24, 21)
    assert a.cigar(24, 21, cg) == "2M6I16M3D"
    p = api.AAProfile.from_bytes(lib, b"TTAAAAAAATTTTTTTTTTTT", 16, 1, -1, -2, -1, -1, -1)
    p.set_gap_close_C(17, -1)
    p.set_gap_close_C(19, 0)
    assert a.align_profile(q, p, (16, 16), 0) == (6, 24, 21)
    assert a.cigar(24, 21, cg) == "2M6I14M3D2M"
    # Trace::cigar from other end positions (scan_block.rs:1469) against the oracle
    ob = ora.Block(100, 100, 16, ora.TRACE)
    op = ora.Profile.from_bytes(b"TTAAAAAAATTTTTTTTTTTT", 16, 1, -1, -2, -1, -1, -1)
    op.set_gap_close_C(17, -1)
    op.set_gap_close_C(19, 0)
    ob.align_profile(ora.Padded(ora.AA, b"TTTTTTTTAAAAAAATTTTTTTTT", 16), op, (16, 16), 0)
    for (i, j) in [(24, 21), (20, 18), (10, 3), (5, 5), (0, 4), (7, 0)]:
        assert a.cigar(i, j, cg) == ob.cigar(i, j)
    # c/example.c examples 1-3 run and agree with the oracle
    a = api.Block(lib, 8, 7, 32, trace=True)
    q, r = api.PaddedBytes(lib, b"AAAAAAAA", 32), api.PaddedBytes(lib, b"AARAAAA", 32)
    res = a.align(q, r, b62, g, (32, 32), 0)
    ob = ora.Block(8, 7, 32, ora.TRACE)
    assert res == ob.align(ora.Padded(ora.AA, b"AAAAAAAA", 32), ora.Padded(ora.AA, b"AARAAAA", 32), ora.AA, ora.builtin("BLOSUM62"), g, (32, 32), 0)
    assert a.cigar(res[1], res[2], api.Cigar(lib, res[1], res[2])) == ob.cigar(res[1], res[2])
