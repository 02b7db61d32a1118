"""The packed 2 x i16 path (block_aligner_b200/csrc/ba_packed.cuh), emulated on the CPU, against the oracle.

Two things are checked beyond plain parity:
* coverage -- the emulated build counts the cells computed by the packed rectangles, by the exact 32-bit
  rectangles and the packed fast steps, so a silent fall-back to the exact path cannot hide a broken packed path;
* the range guard -- scoring schemes whose values leave the i16 range the packed arithmetic is exact on
  (saturating adds in the reference, avx2.rs:26-30) must be routed to the exact path and stay bit-exact.
"""
import ctypes as C

import pytest

import backend
import parity
from block_aligner_b200 import api, workloads

P = workloads.params
NOISY = dict(sub_rate=0.05, ins_rate=0.04, del_rate=0.04, long_indel_mean=1.5, long_indel_len=50.0)


@pytest.fixture(scope="module")
def env():
    lib = backend.emu_lib()
    lib.L.ba_emu_stats.argtypes = [C.c_void_p, C.c_int]
    return lib, api.Aligner(lib)


def stats(lib, reset=True):
    out = (C.c_uint64 * 4)()
    lib.L.ba_emu_stats(out, int(reset))
    return dict(pk_cells=out[0], exact_cells=out[1], fast_steps=out[2], big_cells=out[3])


@pytest.mark.parametrize("flags", [0, api.XDROP])
@pytest.mark.parametrize("size", [(32, 32), (32, 256), (64, 512), (128, 128), (256, 256)])
def test_dna_packed_coverage(env, flags, size):
    lib, al = env
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=size, x_drop=50, flags=flags, stream=51,
             gen=P(alphabet=0, len_dist=0, len_min=600, len_max=2500, suffix_len=150, big_indel_prob=0.5, big_indel_min=60,
                   big_indel_max=250, **NOISY))
    stats(lib)
    assert parity.check_workload(lib, al, w, 10, seed=31 + flags) == 0
    s = stats(lib)
    fast_cells = s["fast_steps"] * 8 * size[0] if size[0] in (32, 64) else 0
    packed = s["pk_cells"] + fast_cells + s["big_cells"]
    if size[1] <= 256:
        assert packed > 4 * s["exact_cells"], s      # the packed path carries the bulk of the work
    else:
        assert packed > 0, s                         # (64, 512): no global-border kernel, rectangles taller than 256 rows take the exact path
    if size[0] in (32, 64):
        assert s["fast_steps"] > 0, s


@pytest.mark.parametrize("flags", [0, api.XDROP])
def test_protein_and_bytes_packed(env, flags):
    lib, al = env
    w = dict(workloads.WORKLOADS["C3_uniclust_protein_global"])
    w["flags"], w["x_drop"] = flags, 40
    stats(lib)
    assert parity.check_workload(lib, al, w, 40, seed=5) == 0
    s = stats(lib)
    assert s["fast_steps"] > 0 and s["pk_cells"] > 0, s
    w = dict(scoring=api.SCORING_BYTE, matrix="BYTES1", gaps=(-2, -1), size=(32, 128), x_drop=30, flags=flags, stream=52,
             gen=P(alphabet=1, len_dist=0, len_min=200, len_max=900, sub_rate=0.1, ins_rate=0.03, del_rate=0.03,
                   long_indel_mean=1.0, long_indel_len=40.0))
    stats(lib)
    assert parity.check_workload(lib, al, w, 16, seed=9) == 0
    assert stats(lib)["fast_steps"] > 0


@pytest.mark.parametrize("flags", [0, api.XDROP])
@pytest.mark.parametrize("matrix,gaps", [((127, -128), (-128, -127)), ((120, -100), (-100, -90)), ((1, -1), (-128, -1)),
                                         ((127, -1), (-3, -2)), ((2, -120), (-120, -100))])
def test_range_guard_extreme_scores(env, flags, matrix, gaps):
    """Scores at the edge of i8: stored values run into the saturation of the reference's i16 adds, where packed
    (wrapping) arithmetic would differ -- those rectangles must take the exact path and every result must match."""
    lib, al = env
    w = dict(scoring=api.SCORING_NUC, matrix=matrix, gaps=gaps, size=(32, 256), x_drop=3000, flags=flags, stream=53,
             gen=P(alphabet=0, len_dist=0, len_min=300, len_max=1800, suffix_len=200, big_indel_prob=0.5, big_indel_min=60,
                   big_indel_max=250, **NOISY))
    stats(lib)
    assert parity.check_workload(lib, al, w, 12, seed=41 + flags) == 0
    s = stats(lib)
    # which of these leave the packed path's exact range depends on the scores: since the first block and the early-break
    # steps of global alignments run on the packed path too, only the schemes whose values really saturate still do
    print(matrix, gaps, flags, s)
    if (matrix, gaps) in (((127, -128), (-128, -127)), ((2, -120), (-120, -100))):
        assert s["exact_cells"] > 0, s


def test_range_guard_long_score_drift(env):
    """Long alignments: the running offset moves by tens of thousands while stored values stay re-centred; fast
    steps must keep running (no spurious guard failures) and agree with the oracle."""
    lib, al = env
    w = dict(scoring=api.SCORING_NUC, matrix=(5, -4), gaps=(-8, -2), size=(32, 128), x_drop=0, flags=0, stream=54,
             gen=P(alphabet=0, len_dist=0, len_min=12000, len_max=14000, sub_rate=0.02, ins_rate=0.01, del_rate=0.01))
    stats(lib)
    assert parity.check_workload(lib, al, w, 3, seed=3) == 0
    s = stats(lib)
    assert s["fast_steps"] * 8 * 32 > 20 * s["exact_cells"], s


@pytest.mark.parametrize("flags", [0, api.XDROP])
@pytest.mark.parametrize("size", [(32, 256), (64, 128)])
def test_profile_fast_phase_coverage(env, flags, size):
    """sequence-to-profile batches without TRACE run their shift steps in the packed fast phase (pkp_cols8)"""
    lib, al = env
    stats(lib)
    assert parity.check_pssm(lib, al, 24, 5 + flags, False, (0, 0), False, size=size, flags=flags) == 0
    s = stats(lib)
    assert s["fast_steps"] * 8 * size[0] > s["exact_cells"], s
    stats(lib)
    assert parity.check_pssm(lib, al, 12, 9, True, (1, 0), True, size=size, flags=flags | api.TRACE) == 0
    assert stats(lib)["fast_steps"] == 0      # TRACE profile batches stay on the exact path


@pytest.mark.parametrize("flags", [0, api.XDROP, api.TRACE, api.TRACE | api.XDROP])
@pytest.mark.parametrize("size", [(32, 1024), (64, 2048), (64, 4096)])
def test_tall_rectangles_chained_chunks(env, flags, size):
    """Rectangles taller than 256 rows on the packed path (place_rect_pk_tall): 256-row chunks chained through the bottom
    row of the chunk above -- D, R (the vertical-gap scan has to continue across the chunk border: big insertions) and the
    trace bit. Kernels with max block >= 1024 and a fast phase (min block 32 / 64) carry that code."""
    lib, al = env
    w = dict(scoring=api.SCORING_NUC, matrix=(2, -4), gaps=(-6, -2), size=size, x_drop=400, flags=flags, stream=52,
             gen=P(alphabet=0, len_dist=0, len_min=2500, len_max=6000, suffix_len=200, big_indel_prob=0.9, big_indel_min=300,
                   big_indel_max=1500, **NOISY))
    stats(lib)
    assert parity.check_workload(lib, al, w, 6, seed=3 + flags) == 0
    s = stats(lib)
    # the grow rectangles of 512+ rows went through the packed path (a 2048-column rectangle leaves the guard only
    # 32767 - 2048 (max score + |open|) of input range: the grow to 4096 mostly stays on the exact path)
    if size[1] <= 2048:
        assert s["pk_cells"] > 4 * s["exact_cells"], s


@pytest.mark.parametrize("flags", [api.XDROP, api.TRACE | api.XDROP])
def test_tall_rectangles_protein(env, flags):
    """the same on the AAMatrix scorer (two LDS.U16 per pair of cells instead of the nucleotide table)"""
    lib, al = env
    w = dict(scoring=api.SCORING_AA, matrix="BLOSUM62", gaps=(-11, -1), size=(32, 1024), x_drop=100, flags=flags, stream=77,
             gen=P(alphabet=1, len_dist=0, len_min=2500, len_max=6000, suffix_len=200, big_indel_prob=0.9, big_indel_min=300,
                   big_indel_max=1500, **NOISY))
    stats(lib)
    assert parity.check_workload(lib, al, w, 6, seed=3 + flags) == 0
    s = stats(lib)
    assert s["pk_cells"] > 4 * s["exact_cells"], s
