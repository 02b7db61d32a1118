"""Threading contract of the drop-in API (SURVEY.md 8b "Threading"; reference: one `Block` per thread,
src/scan_block.rs:1716): Part 1 handles are not thread-safe per handle but safe across handles, although every
legacy call shares one process-wide BaAligner; Part 2 calls on one BaAligner may come from several threads.
ctypes releases the GIL during the foreign call, so these threads really overlap inside the library."""
import ctypes as C
import threading

import numpy as np
import pytest

import backend
from block_aligner_b200 import api


def _legacy_worker(lib, tid, n, out, errs):
    try:
        b62 = C.addressof((C.c_int8 * 864).in_dll(lib.L, "BLOSUM62"))
        rng = np.random.default_rng(100 + tid)
        blk = api.Block(lib, 400, 400, 64, trace=True, x_drop=(tid % 2 == 1))
        cg = api.Cigar(lib, 400, 400)
        res = []
        for _ in range(n):
            ln = int(rng.integers(5, 300))
            r = bytes(rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8), ln))
            q = bytearray(r)
            for p in rng.integers(0, ln, size=ln // 8):
                q[int(p)] = ord("W")
            q = bytes(q[: ln - int(rng.integers(0, 4))])
            pq, pr = api.PaddedBytes(lib, q, 64), api.PaddedBytes(lib, r, 64)
            a = blk.align(pq, pr, b62, (-11, -1), (16, 64), 30)
            res.append((q, r, a, blk.cigar_eq(pq, pr, a[1], a[2], cg)))
        out[tid] = res
    except Exception as e:       # pragma: no cover
        errs.append(repr(e))


def _check_legacy(lib, n_threads, n):
    out, errs = {}, []
    th = [threading.Thread(target=_legacy_worker, args=(lib, t, n, out, errs)) for t in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    # the same work serially, one thread
    ser = {}
    for t in range(n_threads):
        _legacy_worker(lib, t, n, ser, errs)
    assert not errs, errs
    assert out == ser


def _batch_worker(lib, al, tid, out, errs):
    try:
        rng = np.random.default_rng(7 + tid)
        qs, rs = [], []
        for _ in range(40):
            ln = int(rng.integers(20, 600))
            r = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), ln))
            q = bytearray(r)
            for p in rng.integers(0, ln, size=ln // 10):
                q[int(p)] = ord("A")
            qs.append(bytes(q)); rs.append(r)
        got = []
        for _ in range(3):
            res, cig, _ = al.align_batch(qs, rs, api.SCORING_NUC, lib.builtin_matrix("NW1")[1], (-2, -1), (32, 128), 40,
                                         api.TRACE | api.XDROP, True)
            got.append((res, cig))
        assert got[0] == got[1] == got[2]
        out[tid] = got[0]
    except Exception as e:       # pragma: no cover
        errs.append(repr(e))


def _check_batch(lib, n_threads):
    al = api.Aligner(lib, 0)
    out, errs = {}, []
    th = [threading.Thread(target=_batch_worker, args=(lib, al, t, out, errs)) for t in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    ser = {}
    for t in range(n_threads):
        _batch_worker(lib, al, t, ser, errs)
    assert not errs, errs
    assert out == ser


def test_legacy_blocks_one_per_thread_emulated():
    _check_legacy(backend.emu_lib(), 4, 12)


def test_batch_calls_from_several_threads_on_one_aligner_emulated():
    _check_batch(backend.emu_lib(), 4)


@pytest.mark.gpu
def test_legacy_blocks_one_per_thread_gpu():
    _check_legacy(backend.cuda_lib(), 8, 40)


@pytest.mark.gpu
def test_batch_calls_from_several_threads_on_one_aligner_gpu():
    _check_batch(backend.cuda_lib(), 8)
