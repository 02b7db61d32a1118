"""The one-call entry points the benchmark's end-to-end leg times -- ba_align_batch, ba_align_batch_cigar,
ba_align_batch_pssm, ba_align_batch_profiles, ba_align_batch_exp -- on the GPU, against the CPU oracle. They pipeline
chunks on separate CUDA streams through pooled device buffers and pinned staging, which the emulated build (synchronous
memcpy "streams") cannot exercise. Needs a GPU: `pytest -m gpu`."""
import ctypes as C

import numpy as np
import pytest

import backend
import ora
import parity
from block_aligner_b200 import api, workloads

import os

pytestmark = pytest.mark.gpu
P = workloads.params
# development aid: BA_TEST_BACKEND=emu BA_TEST_SCALE=0.01 runs this module's logic on the emulated library (with -m gpu)
EMU = os.environ.get("BA_TEST_BACKEND") == "emu"
SCALE = float(os.environ.get("BA_TEST_SCALE", "1"))


def N(n):
    return max(8, int(n * SCALE))


@pytest.fixture(scope="module")
def env():
    lib = backend.emu_lib() if EMU else backend.cuda_lib()
    return lib, api.Aligner(lib, 0)


def _c2_like(n, lmin, lmax, stream=2, seed=77):
    gen = P(alphabet=0, len_dist=0, len_min=lmin, len_max=lmax, sub_rate=0.04, ins_rate=0.04, del_rate=0.04,
            long_indel_mean=1.0, long_indel_len=60.0, suffix_len=200)
    return workloads.generate(gen, n, seed=seed, stream=stream)


@pytest.mark.parametrize("chunks", ["1", "2", "5", "8"])
@pytest.mark.parametrize("geom", ["0", "1"])
def test_align_batch_pipelined_chunks_against_oracle(env, chunks, geom, monkeypatch):
    lib, al = env
    monkeypatch.setenv("BA_PIPELINE_CHUNKS", chunks)
    monkeypatch.setenv("BA_PIPELINE_GEOM", geom)
    qa, qo, ra, ro = _c2_like(N(6000), 300, 3000)
    m = lib.builtin_matrix("NW1")[1]
    cfg = al.config(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False, qa, qo, ra, ro)
    for rep in range(3):          # repeated calls reuse pooled buffers and streams
        res, st = parity.abi_align_batch(lib, al, cfg, qa, qo, ra, ro)
        assert parity.compare_abi(f"chunks{chunks}/geom{geom}/rep{rep}", res, None, st, exp) == 0
        assert st.n_failed == 0


def test_align_batch_automatic_geometric_split_against_oracle(env):
    """A batch large enough (>= 65536 pairs, >= 512 MB) for the automatic 5-chunk geometric split the C2 e2e number uses."""
    lib, al = env
    if EMU:
        pytest.skip("needs the GPU-sized batch")
    n = 66000
    qa, qo, ra, ro = _c2_like(n, 4000, 4400, seed=5)
    assert int(qo[-1] + ro[-1]) >= (512 << 20)
    m = lib.builtin_matrix("NW1")[1]
    cfg = al.config(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False)
    res, st = parity.abi_align_batch(lib, al, cfg, qa, qo, ra, ro)
    assert st.kernel_launches == 10        # 5 chunks x (convert/pad + align)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False, qa, qo, ra, ro)
    assert parity.compare_abi("auto-geometric", res, None, st, exp) == 0
    lib.check(lib.L.ba_trim(al.h))


def test_align_batch_mid_size_geometric_split_and_staged_copy(env, monkeypatch):
    """>= 16384 pairs and >= 128 MB: four geometric chunks (the shards of a multi-GPU call); the inputs are numpy arrays in
    pageable memory, i.e. they go through the threaded staging copy (h2d_seq)."""
    lib, al = env
    if EMU:
        pytest.skip("needs the GPU-sized batch")
    qa, qo, ra, ro = _c2_like(17000, 3900, 4300, seed=9)
    assert (128 << 20) <= int(qo[-1] + ro[-1]) < (512 << 20)
    m = lib.builtin_matrix("NW1")[1]
    cfg = al.config(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False, qa, qo, ra, ro)
    for stage_threads in ("4", "1", "8"):
        monkeypatch.setenv("BA_STAGE_THREADS", stage_threads)
        res, st = parity.abi_align_batch(lib, al, cfg, qa, qo, ra, ro)
        assert st.kernel_launches == 8         # 4 chunks x (convert/pad + align)
        assert parity.compare_abi(f"mid-geometric/threads{stage_threads}", res, None, st, exp) == 0
    lib.check(lib.L.ba_trim(al.h))


def test_staged_copy_of_small_batches(env, monkeypatch):
    """BA_STAGE_MIN_BYTES=1: every sequence arena in pageable memory takes the staging path, whatever its size"""
    lib, al = env
    monkeypatch.setenv("BA_STAGE_MIN_BYTES", "1")
    monkeypatch.setenv("BA_PIPELINE_CHUNKS", "3")
    qa, qo, ra, ro = _c2_like(N(3000), 300, 3000, seed=21)
    m = lib.builtin_matrix("NW1")[1]
    cfg = al.config(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False, qa, qo, ra, ro)
    for rep in range(2):
        res, st = parity.abi_align_batch(lib, al, cfg, qa, qo, ra, ro)
        assert parity.compare_abi(f"staged-small/rep{rep}", res, None, st, exp) == 0


@pytest.mark.parametrize("chunks", ["1", "2", "3"])
def test_align_batch_cigar_against_oracle(env, chunks, monkeypatch):
    lib, al = env
    monkeypatch.setenv("BA_PIPELINE_CHUNKS", chunks)
    qa, qo, ra, ro = _c2_like(N(1500), 300, 3000, seed=9)
    m = api.nuc_matrix(2, -4)
    flags = api.XDROP | api.TRACE
    cfg = al.config(api.SCORING_NUC, m, (-6, -2), (64, 512), 200, flags, True)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-6, -2), (64, 512), 200, flags, True, qa, qo, ra, ro)
    for rep in range(2):
        res, cigs, st = parity.abi_align_batch_cigar(lib, al, cfg, qa, qo, ra, ro)
        assert parity.compare_abi(f"cigar/chunks{chunks}/rep{rep}", res, cigs, st, exp) == 0


def test_align_batch_cigar_global_protein(env):
    lib, al = env
    w = workloads.WORKLOADS["C3_uniclust_protein_global"]
    qa, qo, ra, ro = workloads.generate(w["gen"], N(4000), seed=3, stream=3)
    m = lib.builtin_matrix("BLOSUM62")[1]
    cfg = al.config(api.SCORING_AA, m, (-11, -1), (32, 256), 0, api.TRACE, False)
    exp = parity.oracle_batch(api.SCORING_AA, m, (-11, -1), (32, 256), 0, api.TRACE, False, qa, qo, ra, ro)
    res, cigs, st = parity.abi_align_batch_cigar(lib, al, cfg, qa, qo, ra, ro)
    assert parity.compare_abi("cigar/protein", res, cigs, st, exp) == 0


def test_align_batch_pssm_and_profiles_against_oracle(env):
    """C4 through both one-call profile entry points (device-built PSSM batch, host AAProfile objects)"""
    lib, al = env
    w = workloads.WORKLOADS["C4_seq_to_profile_xdrop"]
    n = max(N(3000), 400)
    qa, qo, ra, ro = workloads.generate(w["gen"], n, seed=1234, stream=w["stream"])
    cfg = al.config(api.SCORING_PROFILE, None, None, w["size"], w["x_drop"], w["flags"], False)
    oprofs = parity.make_ora_profiles(ra, ro, w["size"][1], -10, -1, 4321)
    exp = parity.oracle_batch(api.SCORING_PROFILE, None, None, w["size"], w["x_drop"], w["flags"], False, qa, qo, ra, ro, profiles=oprofs)
    pb = workloads.make_pssm_batch(lib, ra, ro, seed=4321)
    out = np.zeros(n, dtype=parity.ABI_RES_DT)
    st = api.BaStats()
    for rep in range(2):
        out[:] = 0
        lib.check(lib.L.ba_align_batch_pssm(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, C.byref(pb.c), out.ctypes.data, C.byref(st)))
        res = np.stack([out["score"].astype(np.int64), out["q"].astype(np.int64), out["r"].astype(np.int64)], axis=1)
        assert parity.compare_abi(f"pssm/rep{rep}", res, None, st, exp) == 0
    profs = workloads.make_lib_profiles(lib, ra[:int(ro[400])], ro[:401], w["size"][1], seed=4321)
    arr = (C.c_void_p * 400)(*[p.h for p in profs])
    out[:] = 0
    lib.check(lib.L.ba_align_batch_profiles(al.h, C.byref(cfg), 400, qa.ctypes.data, qo.ctypes.data, arr, out.ctypes.data, C.byref(st)))
    res = np.stack([out["score"][:400].astype(np.int64), out["q"][:400].astype(np.int64), out["r"][:400].astype(np.int64)], axis=1)
    assert (res == exp[0][:400]).all() and int(st.cells) == int(exp[1][:400].sum())


def _exp_loop_oracle(kind, matrix, gaps, size, x_drop, flags, qs, rs, targets):
    """Block::align_exp (scan_block.rs:884-902) on the oracle, pair by pair"""
    out = []
    for q, r, t in zip(qs, rs, targets):
        ob = ora.Block(len(q), len(r), size[1], flags)
        pq, pr = ora.Padded(kind, q, size[1]), ora.Padded(kind, r, size[1])
        mn, used, res = max(size[0], 16), None, None
        while mn <= size[1]:
            res = ob.align(pq, pr, kind, matrix, gaps, (mn, size[1]), x_drop)
            if res[0] >= t:
                used = mn
                break
            mn *= 2
        out.append((res, used))
    return out


@pytest.mark.parametrize("flags", [0, api.XDROP])
def test_align_batch_exp_against_oracle_loop(env, flags):
    lib, al = env
    gen = P(alphabet=0, len_dist=0, len_min=300, len_max=1500, sub_rate=0.05, ins_rate=0.03, del_rate=0.03,
            big_indel_prob=0.9, big_indel_min=40, big_indel_max=200)
    n = N(600)
    qa, qo, ra, ro = workloads.generate(gen, n, stream=51, seed=3 + flags)
    qs = [qa[int(qo[k]):int(qo[k + 1])].tobytes() for k in range(n)]
    rs = [ra[int(ro[k]):int(ro[k + 1])].tobytes() for k in range(n)]
    nw1 = lib.builtin_matrix("NW1")[1]
    full, _, _ = al.align_batch(qs, rs, api.SCORING_NUC, nw1, (-2, -1), (512, 512), 10000, flags)
    targets = [f[0] for f in full]
    targets[7] = 10 ** 6          # unreachable: min_size_used = None, result of the last attempt
    res, used, _ = api.align_batch_exp(al, qs, rs, api.SCORING_NUC, nw1, (-2, -1), (32, 512), targets, x_drop=10000, flags=flags)
    exp = _exp_loop_oracle(ora.NUC, ora.nw1(), (-2, -1), (32, 512), 10000, flags, qs, rs, targets)
    assert [(tuple(r), u) for r, u in zip(res, used)] == [(tuple(e[0]), e[1]) for e in exp]
    assert used[7] is None and len(set(used)) >= (3 if flags == 0 else 2), "test inputs should need retries at several sizes"
