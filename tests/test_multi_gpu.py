"""In-library multi-GPU entry points (ba_align_batch_multi*) and the packed nucleotide input format
(BA_INPUT_NUC4). CPU part: the emulated build with three pretend devices (BA_EMU_DEVICES), i.e. the sharding, the
per-shard threads and the in-order delivery; GPU part: every GPU of the box."""
import ctypes as C
import os

import numpy as np
import pytest

import backend
import parity
from block_aligner_b200 import api, workloads

P = workloads.params


def _batch(n, seed=5, lmin=100, lmax=1200):
    gen = P(alphabet=0, len_dist=0, len_min=lmin, len_max=lmax, sub_rate=0.04, ins_rate=0.04, del_rate=0.04,
            long_indel_mean=1.0, long_indel_len=40.0, suffix_len=80)
    return workloads.generate(gen, n, seed=seed, stream=2)


def _multi(lib, devices, n_dev, cfg, qa, qo, ra, ro):
    n = len(qo) - 1
    out = np.zeros(max(n, 1), dtype=parity.ABI_RES_DT)
    st = api.BaStats()
    dv = (C.c_int * len(devices))(*devices) if devices is not None else None
    lib.check(lib.L.ba_align_batch_multi(dv, n_dev, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data,
                                         out.ctypes.data, C.byref(st)))
    return np.stack([out["score"][:n].astype(np.int64), out["q"][:n].astype(np.int64), out["r"][:n].astype(np.int64)], axis=1), st


def _multi_cigar(lib, n_dev, cfg, qa, qo, ra, ro, cap=None):
    n = len(qo) - 1
    out = np.zeros(max(n, 1), dtype=parity.ABI_RES_DT)
    cap = cap or int(qo[-1] + ro[-1]) + 5 * n + 8
    runs = np.zeros(cap, dtype=np.uint32)
    off = np.zeros(max(n, 1), dtype=np.uint64)
    ln = np.zeros(max(n, 1), dtype=np.uint32)
    used = C.c_size_t()
    st = api.BaStats()
    lib.check(lib.L.ba_align_batch_multi_cigar(None, n_dev, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data,
                                               out.ctypes.data, runs.ctypes.data, cap, off.ctypes.data, ln.ctypes.data, C.byref(used), C.byref(st)))
    assert used.value == int(ln[:n].sum())
    cigs = [runs[int(off[k]):int(off[k]) + int(ln[k])].astype(np.uint64) for k in range(n)]
    return np.stack([out["score"][:n].astype(np.int64), out["q"][:n].astype(np.int64), out["r"][:n].astype(np.int64)], axis=1), cigs, st


def _check_multi(lib, al, n_dev, n, scale=1):
    m = lib.builtin_matrix("NW1")[1]
    qa, qo, ra, ro = _batch(n)
    cfg = al.config(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, api.XDROP, False, qa, qo, ra, ro)
    res, st = _multi(lib, None, n_dev, cfg, qa, qo, ra, ro)
    assert parity.compare_abi("multi", res, None, st, exp) == 0 and st.n_failed == 0
    if n_dev >= 2:     # explicit device list, reversed
        res, st = _multi(lib, list(range(n_dev))[::-1], n_dev, cfg, qa, qo, ra, ro)
        assert parity.compare_abi("multi-rev", res, None, st, exp) == 0
    # CIGARs: absolute run offsets into one caller buffer
    flags = api.XDROP | api.TRACE
    cfg = al.config(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, flags, True)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-2, -1), (32, 256), 50, flags, True, qa, qo, ra, ro)
    res, cigs, st = _multi_cigar(lib, n_dev, cfg, qa, qo, ra, ro)
    assert parity.compare_abi("multi-cigar", res, cigs, st, exp) == 0
    # a buffer below the worst case is shared out proportionally (still enough here: CIGARs are far shorter than |q| + |r|)
    res, cigs, st = _multi_cigar(lib, n_dev, cfg, qa, qo, ra, ro, cap=int(qo[-1] + ro[-1]) // 2)
    assert parity.compare_abi("multi-cigar-small", res, cigs, st, exp) == 0
    # profiles
    w = workloads.WORKLOADS["C4_seq_to_profile_xdrop"]
    npf = max(8, n // 2)
    qa, qo, ra, ro = workloads.generate(w["gen"], npf, seed=12, stream=4)
    cfg = al.config(api.SCORING_PROFILE, None, None, (32, 128), w["x_drop"], w["flags"], False)
    pb = workloads.make_pssm_batch(lib, ra, ro, seed=31)
    exp = parity.oracle_batch(api.SCORING_PROFILE, None, None, (32, 128), w["x_drop"], w["flags"], False, qa, qo, ra, ro,
                              profiles=parity.make_ora_profiles(ra, ro, 128, -10, -1, 31))
    out = np.zeros(npf, dtype=parity.ABI_RES_DT)
    st = api.BaStats()
    lib.check(lib.L.ba_align_batch_multi_pssm(None, n_dev, C.byref(cfg), npf, qa.ctypes.data, qo.ctypes.data, C.byref(pb.c),
                                              out.ctypes.data, C.byref(st)))
    res = np.stack([out["score"].astype(np.int64), out["q"].astype(np.int64), out["r"].astype(np.int64)], axis=1)
    assert parity.compare_abi("multi-pssm", res, None, st, exp) == 0


def test_multi_three_emulated_devices(monkeypatch):
    monkeypatch.setenv("BA_EMU_DEVICES", "3")
    lib = backend.emu_lib()
    assert lib.L.ba_device_count() == 3
    al = api.Aligner(lib)
    _check_multi(lib, al, 3, 40)
    _check_multi(lib, al, 0, 13)        # n_dev <= 0: every device
    # fewer pairs than devices, and an empty batch
    m = lib.builtin_matrix("NW1")[1]
    cfg = al.config(api.SCORING_NUC, m, (-2, -1), (32, 64), 0, 0, False)
    qa, qo, ra, ro = _batch(2)
    res, st = _multi(lib, None, 3, cfg, qa, qo, ra, ro)
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-2, -1), (32, 64), 0, 0, False, qa, qo, ra, ro)
    assert parity.compare_abi("multi-2", res, None, st, exp) == 0
    z = np.zeros(1, dtype=np.uint64)
    res, st = _multi(lib, None, 3, cfg, qa, z, ra, z)
    assert len(res) == 0 and st.cells == 0
    with pytest.raises(api.BlockAlignerError, match="twice"):
        _multi(lib, [1, 1], 2, cfg, qa, qo, ra, ro)
    # an error in one shard (bad character) is reported with its device
    bad = qa.copy()
    bad[int(qo[1]) + 3] = ord("-")
    with pytest.raises(api.BlockAlignerError, match="device"):
        _multi(lib, None, 2, cfg, bad, qo, ra, ro)
    lib.L.ba_multi_release()


def _check_nuc4(lib, al, n):
    m = api.nuc_matrix(2, -4)
    qa, qo, ra, ro = _batch(n, seed=9)
    # sprinkle N and lower case
    rng = np.random.default_rng(1)
    qa = qa.copy()
    qa[rng.integers(0, len(qa), size=len(qa) // 50)] = ord("N")
    low = rng.integers(0, len(qa), size=len(qa) // 20)
    qa[low] = qa[low] | 32
    flags = api.XDROP | api.TRACE
    exp = parity.oracle_batch(api.SCORING_NUC, m, (-6, -2), (32, 256), 100, flags, True, qa, qo, ra, ro)
    pq, pqo = api.pack_nuc4(lib, qa, qo)
    pr, pro = api.pack_nuc4(lib, ra, ro)
    assert pq.nbytes <= qa.nbytes // 2 + 2
    cfg = al.config(api.SCORING_NUC, m, (-6, -2), (32, 256), 100, flags | api.INPUT_NUC4, True)
    res, cigs, st = parity.abi_align_batch_cigar(lib, al, cfg, pq, pqo, pr, pro)
    assert parity.compare_abi("nuc4", res, cigs, st, exp) == 0
    # a sub-range that starts on an odd nibble, and the upload / run / download path
    lo = next(k for k in range(1, n) if int(qo[k]) & 1 and int(ro[k]) & 1) if n > 4 else 0
    b = al.upload(cfg, pq, pqo[lo:], pr, pro[lo:])
    b.run()
    r = b.download()
    assert (r["score"] == exp[0][lo:, 0]).all() and (r["query_idx"] == exp[0][lo:, 1]).all()
    assert [b.cigar_string(k) for k in range(n - lo)] == [api.runs_to_string(c) for c in exp[2][lo:]]
    b.free()
    # reversed on the device from packed input == oracle on host-reversed ASCII
    def rev(a, off):
        o = a.copy()
        for k in range(len(off) - 1):
            o[int(off[k]):int(off[k + 1])] = a[int(off[k]):int(off[k + 1])][::-1]
        return o
    cfg = al.config(api.SCORING_NUC, m, (-6, -2), (32, 256), 100, api.XDROP | api.INPUT_NUC4 | api.REV_QUERY | api.REV_REFERENCE, False)
    res, st = parity.abi_align_batch(lib, al, cfg, pq, pqo, pr, pro)
    exp_r = parity.oracle_batch(api.SCORING_NUC, m, (-6, -2), (32, 256), 100, api.XDROP, False, rev(qa, qo), qo, rev(ra, ro), ro)
    assert parity.compare_abi("nuc4-rev", res, None, st, exp_r) == 0
    # code 0 ('=') is not a base
    bad = pq.copy()
    bad[3] &= 0x0f
    with pytest.raises(api.BlockAlignerError, match="alphabet"):
        parity.abi_align_batch(lib, al, cfg, bad, pqo, pr, pro)
    with pytest.raises(api.BlockAlignerError, match="NUC4"):
        parity.abi_align_batch(lib, al, al.config(api.SCORING_AA, lib.builtin_matrix("BLOSUM62")[1], (-11, -1), (32, 32), 0,
                                                  api.INPUT_NUC4, False), pq, pqo, pr, pro)


def test_nuc4_input_emulated(monkeypatch):
    lib = backend.emu_lib()
    al = api.Aligner(lib)
    _check_nuc4(lib, al, 30)
    monkeypatch.setenv("BA_PIPELINE_CHUNKS", "3")      # chunk boundaries at odd nibbles
    _check_nuc4(lib, al, 30)


def test_pack_nuc4_helper():
    lib = backend.emu_lib()
    s = np.frombuffer(b"ACGTNacgtnMRWSYKVHDB", dtype=np.uint8)
    out = np.full(12, 0xFF, dtype=np.uint8)
    assert lib.L.ba_pack_nuc4(s.ctypes.data, len(s), out.ctypes.data, 1) == 0
    code = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    nibs = [(int(out[i >> 1]) >> 4) if i % 2 == 0 else (int(out[i >> 1]) & 15) for i in range(22)]
    assert nibs[0] == 15 and nibs[21] == 15                       # neighbours untouched
    assert nibs[1:21] == [code[chr(c).upper()] for c in s]
    assert lib.L.ba_pack_nuc4(np.frombuffer(b"AC-T", dtype=np.uint8).ctypes.data, 4, out.ctypes.data, 0) == 3


@pytest.mark.gpu
def test_multi_all_gpus_of_the_box():
    lib = backend.cuda_lib()
    n_dev = lib.L.ba_device_count()
    assert n_dev >= 1
    al = api.Aligner(lib, 0)
    _check_multi(lib, al, n_dev, 3000)
    if n_dev >= 2:
        _check_multi(lib, al, 2, 500)
    lib.L.ba_multi_release()


@pytest.mark.gpu
def test_nuc4_input_gpu(monkeypatch):
    lib = backend.cuda_lib()
    al = api.Aligner(lib, 0)
    _check_nuc4(lib, al, 2000)
    monkeypatch.setenv("BA_PIPELINE_CHUNKS", "5")
    _check_nuc4(lib, al, 2000)
