"""Parity tests proper: the CUDA product (through its C ABI) against the CPU oracle on identical seeded
inputs, plus the reference's golden vectors. Bit-exact bar for score, query_idx, reference_idx, CIGAR
runs and the path-determined cell count. Needs a GPU: `pytest -m gpu`."""
import numpy as np
import pytest

import backend
import golden_cases as G
import parity
from block_aligner_b200 import api, workloads

pytestmark = pytest.mark.gpu
P = workloads.params
N_MIN64 = 400
NOISY = dict(sub_rate=0.05, ins_rate=0.04, del_rate=0.04, long_indel_mean=1.5, long_indel_len=50.0)


@pytest.fixture(scope="module")
def env():
    lib = backend.cuda_lib()
    return lib, api.Aligner(lib, 0)


def test_golden_batch_api(env):
    G.run_batch_golden(*env)


def test_golden_legacy_api(env):
    G.run_legacy_golden(env[0])


@pytest.mark.parametrize("flags", [0, api.XDROP, api.TRACE, api.TRACE | api.XDROP])
@pytest.mark.parametrize("size", [(16, 16), (16, 64), (32, 32), (32, 256), (64, 64), (128, 512), (256, 2048), (32, 1024)])
def test_dna(env, flags, size):
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=size, x_drop=50, flags=flags, stream=11,
             gen=P(alphabet=0, len_dist=0, len_min=300, len_max=3000, suffix_len=150, **NOISY))
    assert parity.check_workload(*env, w, 600, seed=7 + flags) == 0


@pytest.mark.parametrize("flags", [0, api.XDROP, api.TRACE, api.TRACE | api.XDROP])
@pytest.mark.parametrize("size", [(64, 64), (64, 256), (64, 2048)])
def test_dna_min64(env, flags, size):
    """min block 64: the fast phase with 8 rows per lane (C5 shape)"""
    w = dict(scoring=api.SCORING_NUC, matrix=(2, -4), gaps=(-6, -2), size=size, x_drop=100, flags=flags, stream=21,
             gen=P(alphabet=0, len_dist=0, len_min=300, len_max=2500, suffix_len=150, big_indel_prob=0.5, big_indel_min=80,
                   big_indel_max=300, **NOISY))
    assert parity.check_workload(*env, w, N_MIN64, seed=11 + flags) == 0


@pytest.mark.parametrize("flags", [0, api.XDROP, api.TRACE, api.TRACE | api.XDROP])
@pytest.mark.parametrize("size", [(32, 1024), (64, 2048), (64, 8192)])
def test_tall_rectangles_chained_chunks(env, flags, size):
    """big indels grow the block past 256 rows: packed rectangles in chained 256-row chunks (place_rect_pk_tall)"""
    w = dict(scoring=api.SCORING_NUC, matrix=(2, -4), gaps=(-6, -2), size=size, x_drop=400, flags=flags, stream=52,
             gen=P(alphabet=0, len_dist=0, len_min=2500, len_max=9000, suffix_len=200, big_indel_prob=0.9, big_indel_min=300,
                   big_indel_max=3000, **NOISY))
    assert parity.check_workload(*env, w, 150, seed=3 + flags) == 0


@pytest.mark.parametrize("flags", [0, api.XDROP, api.TRACE, api.TRACE | api.XDROP])
def test_dna_big_indels_force_growth(env, flags):
    w = dict(scoring=api.SCORING_NUC, matrix=(2, -4), gaps=(-6, -2), size=(32, 512), x_drop=200, flags=flags, stream=12,
             gen=P(alphabet=0, len_dist=0, len_min=1500, len_max=4000, suffix_len=100, big_indel_prob=0.8,
                   big_indel_min=100, big_indel_max=400, **NOISY))
    assert parity.check_workload(*env, w, 400) == 0


@pytest.mark.parametrize("flags", [0, api.XDROP, api.TRACE, api.TRACE | api.XDROP])
@pytest.mark.parametrize("size", [(16, 128), (32, 256), (32, 32)])
def test_protein(env, flags, size):
    w = dict(workloads.WORKLOADS["C3_uniclust_protein_global"])
    w["flags"], w["x_drop"], w["size"] = flags, 30, size
    assert parity.check_workload(*env, w, 3000, seed=3) == 0


@pytest.mark.parametrize("flags", [0, api.XDROP, api.TRACE, api.TRACE | api.XDROP])
@pytest.mark.parametrize("size", [(16, 16), (32, 256), (64, 512)])
def test_profile(env, flags, size):
    w = dict(workloads.WORKLOADS["C4_seq_to_profile_xdrop"])
    w["flags"] = flags
    assert parity.check_workload(*env, w, 300, size=size, seed=5) == 0


@pytest.mark.parametrize("flags", [0, api.TRACE, api.TRACE | api.XDROP])
def test_tiny_and_empty(env, flags):
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=(32, 128), x_drop=20, flags=flags, stream=13,
             gen=P(alphabet=0, len_dist=0, len_min=0, len_max=40, sub_rate=0.1, ins_rate=0.1, del_rate=0.1))
    assert parity.check_workload(*env, w, 2000) == 0


def test_byte_matrix(env):
    w = dict(scoring=api.SCORING_BYTE, matrix="BYTES1", gaps=(-2, -1), size=(32, 64), x_drop=0, flags=api.TRACE, stream=14,
             gen=P(alphabet=1, len_dist=0, len_min=50, len_max=300, sub_rate=0.1, ins_rate=0.03, del_rate=0.03))
    assert parity.check_workload(*env, w, 500) == 0


# ---- the BASELINE.json workloads ---------------------------------------------------------------------
def test_C1_full(env):
    w = workloads.WORKLOADS["C1_rand_scan_dna1k"]
    assert parity.check_workload(*env, w, w["n"]) == 0


def test_C2_sample_full_length(env):
    assert parity.check_workload(*env, workloads.WORKLOADS["C2_nanopore_xdrop_10k"], 10000) == 0


def test_C2_with_trace(env):
    assert parity.check_workload(*env, workloads.WORKLOADS["C2_nanopore_xdrop_10k"], 300, first=5000, with_trace=True) == 0


def test_C3_sample(env):
    assert parity.check_workload(*env, workloads.WORKLOADS["C3_uniclust_protein_global"], 100000) == 0


def test_C4_sample(env):
    assert parity.check_workload(*env, workloads.WORKLOADS["C4_seq_to_profile_xdrop"], 1500) == 0


def test_C4_large_sample_device_built_profiles(env):
    """20 000 C4 pairs through the PSSM batch (profiles built on the device) against the oracle"""
    lib, al = env
    w = workloads.WORKLOADS["C4_seq_to_profile_xdrop"]
    n = 20000
    qa, qo, ra, ro = workloads.generate(w["gen"], n, seed=1234, stream=w["stream"])
    pb = workloads.make_pssm_batch(lib, ra, ro, seed=11)
    got = parity.run_lib(lib, al, api.SCORING_PROFILE, None, None, w["size"], w["x_drop"], w["flags"], False, qa, qo, ra, ro, pb)
    op = parity.make_ora_profiles(ra, ro, w["size"][1], -10, -1, 11)
    exp = parity.oracle_batch(api.SCORING_PROFILE, None, None, w["size"], w["x_drop"], w["flags"], False, qa, qo, ra, ro, op)
    assert parity.compare("C4-20k", got, exp) == 0


def test_C5_sample_long_reads_with_trace(env):
    assert parity.check_workload(*env, workloads.WORKLOADS["C5_longread_trace_50k"], 512) == 0


def test_results_independent_of_batch_order_and_reuse(env):
    """Block reuse invariance (scan_block.rs:795-797): same pairs, shuffled and re-run on the same aligner."""
    lib, al = env
    w = workloads.WORKLOADS["C2_nanopore_xdrop_10k"]
    qa, qo, ra, ro = workloads.generate(w["gen"], 400, stream=w["stream"])
    m = workloads.matrix_of(lib, w)
    r1 = parity.run_lib(lib, al, w["scoring"], m, w["gaps"], w["size"], w["x_drop"], w["flags"], False, qa, qo, ra, ro)
    r2 = parity.run_lib(lib, al, w["scoring"], m, w["gaps"], w["size"], w["x_drop"], w["flags"], False, qa, qo, ra, ro)
    assert (r1[0] == r2[0]).all() and (r1[1] == r2[1]).all()


@pytest.mark.parametrize("pool", [True, False])
def test_trace_arena_overflow_is_retried(env, pool, monkeypatch):
    """Unrelated sequences keep the block at its maximum size, which overflows the (deliberately small) per-slot
    trace arenas. With the batch-wide overflow pool the rectangles spill there (one launch); without it those pairs
    are re-run with worst-case arenas (second launch). Bit-exact either way."""
    lib, al = env
    if not pool:
        monkeypatch.setenv("BA_NO_TRACE_POOL", "1")
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=(32, 512), x_drop=0, flags=api.TRACE, stream=31,
             gen=P(alphabet=0, len_dist=0, len_min=1500, len_max=2500, sub_rate=0.75, ins_rate=0.0, del_rate=0.0))
    qa, qo, ra, ro = workloads.generate(w["gen"], 200, stream=31)
    m = workloads.matrix_of(lib, w)
    got = parity.run_lib(lib, al, w["scoring"], m, w["gaps"], w["size"], 0, api.TRACE, True, qa, qo, ra, ro)
    nl = got[3].kernel_launches       # an alignment and a traceback kernel per pass; the retry pass may take several waves
    assert (nl == 2) if pool else (nl >= 4 and nl % 2 == 0), "expected the overflow pool / the retry pass to be used"
    exp = parity.oracle_batch(w["scoring"], m, w["gaps"], w["size"], 0, api.TRACE, True, qa, qo, ra, ro)
    assert parity.compare("overflow-retry", got, exp) == 0


@pytest.mark.parametrize("mode", [api.LOCAL_START, api.FREE_QUERY_START_GAPS])
@pytest.mark.parametrize("flags", [0, api.XDROP, api.TRACE, api.TRACE | api.XDROP])
@pytest.mark.parametrize("size", [(32, 32), (32, 256), (64, 512)])
def test_local_start_and_free_query_start_gaps(env, mode, flags, size):
    """Block<_, _, LOCAL_START> / Block<_, _, false, FREE_QUERY_START_GAPS> (scan_block.rs:1130-1136, 1597-1612)"""
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=size, x_drop=60, flags=flags | mode, stream=41,
             gen=P(alphabet=0, len_dist=0, len_min=200, len_max=1500, suffix_len=120, **NOISY))
    assert parity.check_workload(*env, w, 300, seed=17 + flags) == 0


def test_extended_mode_golden_vectors(env):
    """scan_block.rs:2171-2229"""
    lib, al = env
    nw1 = lib.builtin_matrix("NW1")[1]
    T, X, L, F = api.TRACE, api.XDROP, api.LOCAL_START, api.FREE_QUERY_START_GAPS
    for fl, q, r, xd, exp, cig in [
            (T | L, b"CCCCCCCCCCAAAAAA", b"TTTTAAAAAA", 0, (6, 16, 10), "6="),
            (T | X | L, b"CCCCCCCCCCAAAAAACCCCCCCCCCCC", b"TTTTAAAAAATTTTTTT", 100, (6, 16, 10), "6="),
            (T | F, b"AAAAAA", b"CCCCCCCCCCAAAAAA", 0, (6, 6, 16), "6="),
            (T | F, b"AAAAAA", b"CCCCCCCCCCAAATAA", 0, (4, 6, 16), "3=1X2=")]:
        res, cigs, _ = al.align_batch([q], [r], api.SCORING_NUC, nw1, (-2, -1), (32, 32), xd, fl, True)
        assert res[0] == exp and cigs[0] == cig
    E = api.FREE_QUERY_END_GAPS
    for q, r, exp, cig in [(b"AAAAAA", b"AAAAAACCCCCCCCCC", (6, 6, 6), "6="), (b"AAAAAA", b"AAATAACCCCCCCCCC", (4, 6, 6), "3=1X2=")]:
        res, cigs, _ = al.align_batch([q], [r], api.SCORING_NUC, nw1, (-2, -1), (32, 32), 0, T | E, True)
        assert res[0] == exp and cigs[0] == cig
    with pytest.raises(api.BlockAlignerError, match="larger than the query"):
        al.align_batch([b"A" * 40], [b"AAAA"], api.SCORING_NUC, nw1, (-2, -1), (32, 32), 0, E, False)
    with pytest.raises(api.BlockAlignerError, match="X_DROP and FREE_QUERY_END_GAPS"):
        al.align_batch([b"AAAA"], [b"AAAA"], api.SCORING_NUC, nw1, (-2, -1), (32, 32), 0, E | X, False)
    with pytest.raises(api.BlockAlignerError, match="both"):
        al.align_batch([b"AAAA"], [b"AAAA"], api.SCORING_NUC, nw1, (-2, -1), (32, 32), 0, L | F, False)


@pytest.mark.parametrize("mode", [api.LOCAL_START, api.FREE_QUERY_START_GAPS])
def test_extended_modes_protein_and_profile(env, mode):
    w = dict(workloads.WORKLOADS["C3_uniclust_protein_global"])
    w["flags"], w["x_drop"] = api.TRACE | api.XDROP | mode, 40
    assert parity.check_workload(*env, w, 300, seed=23) == 0
    w = dict(workloads.WORKLOADS["C4_seq_to_profile_xdrop"])
    w["flags"] = api.TRACE | mode
    assert parity.check_workload(*env, w, 150, size=(32, 128), seed=29) == 0


@pytest.mark.parametrize("flags", [0, api.TRACE, api.TRACE | api.LOCAL_START, api.FREE_QUERY_START_GAPS])
@pytest.mark.parametrize("size", [(32, 32), (64, 64), (64, 256), (128, 128), (256, 256)])
def test_free_query_end_gaps(env, flags, size):
    """Block<_, false, _, _, FREE_QUERY_END_GAPS> (scan_block.rs:333-368, 1194-1201) on short queries"""
    from test_emu_parity import _short_query_batch
    lib, al = env
    fl = flags | api.FREE_QUERY_END_GAPS
    qa, qo, ra, ro = _short_query_batch(400, size[0] - 1, 40, 900, seed=size[0] + flags)
    m = lib.builtin_matrix("NW1")[1]
    got = parity.run_lib(lib, al, api.SCORING_NUC, m, (-2, -1), size, 0, fl, bool(fl & api.TRACE), qa, qo, ra, ro)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-2, -1), size, 0, fl, bool(fl & api.TRACE), qa, qo, ra, ro)
    assert parity.compare(f"fqe/{flags}/{size}", got, exp) == 0


# ---- profiles built on the device from raw PSSM rows (ba_batch_upload_pssm) ----
@pytest.mark.gpu
@pytest.mark.parametrize("rev,shifts,uniform", [(False, (0, 0), False), (True, (0, 0), False), (False, (1, 0), True),
                                                (True, (0, 1), True)])
def test_pssm_batch_device_profiles(env, rev, shifts, uniform):
    assert parity.check_pssm(*env, 300, 77, rev, shifts, uniform, size=(32, 256)) == 0


@pytest.mark.gpu
def test_pssm_batch_trace(env):
    assert parity.check_pssm(*env, 200, 78, False, (0, 0), False, flags=api.XDROP | api.TRACE) == 0


@pytest.mark.gpu
def test_pssm_batch_matches_host_profiles_c4(env):
    """C4 through both profile entry points: host AAProfile objects (compact upload) and device-built PSSM batch"""
    lib, al = env
    w = workloads.WORKLOADS["C4_seq_to_profile_xdrop"]
    n = 2000
    qa, qo, ra, ro = workloads.generate(w["gen"], n, stream=w["stream"])
    cfg = al.config(api.SCORING_PROFILE, None, None, w["size"], w["x_drop"], w["flags"], False)
    profs = workloads.make_lib_profiles(lib, ra, ro, w["size"][1], seed=99)
    pb = workloads.make_pssm_batch(lib, ra, ro, seed=99)
    outs = []
    for p in (profs, pb):
        b = al.upload(cfg, qa, qo, None, None, p)
        b.run()
        outs.append(b.download().copy())
        b.free()
    assert (outs[0] == outs[1]).all()


# ---- size-independent properties at the BASELINE lengths (tests/test_properties.py) ----
@pytest.mark.gpu
@pytest.mark.parametrize("name,n,scorer", [("C2_nanopore_xdrop_10k", 24, (1, -1)), ("C5_longread_trace_50k", 6, (2, -4))])
def test_cigar_consumes_end_position_and_rescores_full_length(env, name, n, scorer):
    import test_properties as tp
    w = workloads.WORKLOADS[name]
    flags = w["flags"] | api.TRACE
    assert tp.check_cigars(*env, w, n, 1234, w["size"], flags, w["x_drop"], tp.nuc_scorer(*scorer)) == [0, 0]


@pytest.mark.gpu
def test_batch_with_overflow_retry_can_run_again(env, monkeypatch):
    """ba_batch_run twice on a resident batch whose first run needed the retry pass: the retry scratch must not
    replace the first-pass arenas (found on the GPU with C5: illegal address on the second run)."""
    lib, al = env
    monkeypatch.setenv("BA_NO_TRACE_POOL", "1")
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=(32, 512), x_drop=0, flags=api.TRACE, stream=31,
             gen=P(alphabet=0, len_dist=0, len_min=1500, len_max=2500, sub_rate=0.75, ins_rate=0.0, del_rate=0.0))
    qa, qo, ra, ro = workloads.generate(w["gen"], 200, stream=31)
    cfg = al.config(w["scoring"], workloads.matrix_of(lib, w), w["gaps"], w["size"], 0, api.TRACE, True)
    b = al.upload(cfg, qa, qo, ra, ro)
    outs = []
    for _ in range(3):
        st = b.run()
        assert st.kernel_launches >= 4      # first pass + retry pass(es), an alignment and a traceback kernel each
        r = b.download()
        outs.append((r.copy(), [b.cigar_string(k) for k in range(len(qo) - 1)]))
    b.free()
    assert all((o[0] == outs[0][0]).all() and o[1] == outs[0][1] for o in outs[1:])


@pytest.mark.gpu
def test_trace_pool_exhaustion_falls_back_to_retry(env, monkeypatch):
    """A pool too small for the spill: the alignments that cannot get pool space are re-run by the retry pass."""
    lib, al = env
    monkeypatch.setenv("BA_TRACE_POOL_BYTES", "65536")
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=(32, 512), x_drop=0, flags=api.TRACE, stream=31,
             gen=P(alphabet=0, len_dist=0, len_min=1500, len_max=2500, sub_rate=0.75, ins_rate=0.0, del_rate=0.0))
    qa, qo, ra, ro = workloads.generate(w["gen"], 200, stream=31)
    m = workloads.matrix_of(lib, w)
    got = parity.run_lib(lib, al, w["scoring"], m, w["gaps"], w["size"], 0, api.TRACE, True, qa, qo, ra, ro)
    assert got[3].kernel_launches >= 4
    exp = parity.oracle_batch(w["scoring"], m, w["gaps"], w["size"], 0, api.TRACE, True, qa, qo, ra, ro)
    assert parity.compare("pool-exhausted", got, exp) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("rev", [api.REV_QUERY, api.REV_REFERENCE, api.REV_QUERY | api.REV_REFERENCE])
def test_reverse_on_device_matches_host_reversed_inputs(env, rev):
    assert parity.check_reversed(*env, 300, rev) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(32, 256), (64, 2048)])
def test_full_slot_arena_parks_fast_phase_and_spills_to_pool(env, size, monkeypatch):
    """Slot arenas far too small for the path: the fast phase parks the group when its arena cannot take another step
    and the generic phase continues with rectangles from the overflow pool. Bit-exact, no retry launch."""
    lib, al = env
    monkeypatch.setenv("BA_TRACE_ARENA_WORDS", "4096")
    w = dict(scoring=api.SCORING_NUC, matrix="NW1", gaps=(-2, -1), size=size, x_drop=80, flags=api.TRACE | api.XDROP, stream=51,
             gen=P(alphabet=0, len_dist=0, len_min=300, len_max=3000, suffix_len=100, **NOISY))
    qa, qo, ra, ro = workloads.generate(w["gen"], 300, stream=51)
    m = workloads.matrix_of(lib, w)
    got = parity.run_lib(lib, al, w["scoring"], m, w["gaps"], size, 80, w["flags"], True, qa, qo, ra, ro)
    assert got[3].kernel_launches == 2
    exp = parity.oracle_batch(w["scoring"], m, w["gaps"], size, 80, w["flags"], True, qa, qo, ra, ro)
    assert parity.compare("tiny-arena", got, exp) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(32, 256), (64, 2048)])
def test_staged_batch_traceback(env, size, monkeypatch):
    """BA_TB_STAGED=1: the opt-in batch traceback with the rectangles' trace words staged in shared memory by asynchronous
    copies (eight lanes per walk); also with few walks in flight, so that every group of lanes takes several walks."""
    lib, al = env
    monkeypatch.setenv("BA_TB_STAGED", "1")
    w = dict(scoring=api.SCORING_NUC, matrix=(2, -4), gaps=(-6, -2), size=size, x_drop=100, flags=api.TRACE | api.XDROP, stream=57,
             gen=P(alphabet=0, len_dist=0, len_min=300, len_max=4000, suffix_len=150, big_indel_prob=0.5, big_indel_min=80,
                   big_indel_max=300, **NOISY))
    assert parity.check_workload(lib, al, w, 400, seed=5) == 0
    monkeypatch.setenv("BA_TB_WALKS", "64")
    assert parity.check_workload(lib, al, w, 300, seed=6) == 0
