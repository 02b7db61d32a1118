"""ctypes binding of the CPU oracle (oracle/libba_oracle.so) -- test infrastructure only.

Mirrors the reference's Rust surface closely enough that the golden tests below read like
`/root/reference/src/scan_block.rs:1908-2230`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

NUC, AA, BYTE = 0, 1, 2
TRACE, XDROP, LOCAL_START, FREE_QUERY_START_GAPS, FREE_QUERY_END_GAPS = 1, 2, 4, 8, 16
OPS = {0: "?", 1: "M", 2: "=", 3: "X", 4: "I", 5: "D"}


class Result(C.Structure):
    _fields_ = [("score", C.c_int32), ("query_idx", C.c_size_t), ("reference_idx", C.c_size_t)]

    def tup(self):
        return (self.score, self.query_idx, self.reference_idx)


class OpLen(C.Structure):
    _fields_ = [("op", C.c_uint8), ("len", C.c_size_t)]


class Step(C.Structure):
    _fields_ = [("dir", C.c_int32), ("i", C.c_uint32), ("j", C.c_uint32), ("block_size", C.c_uint32),
                ("off", C.c_int32), ("max", C.c_int16), ("right_max", C.c_int16), ("down_max", C.c_int16)]


class Batch(C.Structure):
    _fields_ = [("n", C.c_size_t), ("arena", C.c_void_p),
                ("q_off", C.c_void_p), ("q_len", C.c_void_p), ("r_off", C.c_void_p), ("r_len", C.c_void_p),
                ("profiles", C.c_void_p), ("matrix_kind", C.c_int), ("matrix", C.c_void_p),
                ("gap_open", C.c_int8), ("gap_extend", C.c_int8),
                ("min_size", C.c_size_t), ("max_size", C.c_size_t), ("x_drop", C.c_int32),
                ("flags", C.c_int), ("cigar_eq", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ORACLE_DIR, "libba_oracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.ora_block_new.restype = C.c_void_p
    L.ora_block_new.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
    L.ora_block_free.argtypes = [C.c_void_p]
    L.ora_align.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                            C.c_int8, C.c_int8, C.c_size_t, C.c_size_t, C.c_int32]
    L.ora_align_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int32]
    L.ora_res.restype = Result
    L.ora_res.argtypes = [C.c_void_p]
    L.ora_cells.restype = C.c_uint64
    L.ora_cells.argtypes = [C.c_void_p]
    L.ora_steps.restype = C.c_uint64
    L.ora_steps.argtypes = [C.c_void_p]
    L.ora_cigar.restype = C.c_long
    L.ora_cigar.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                            C.c_size_t, C.c_void_p, C.c_size_t]
    L.ora_enable_step_log.argtypes = [C.c_void_p, C.c_int]
    L.ora_align_logged_steps.restype = C.c_size_t
    L.ora_align_logged_steps.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.ora_profile_new.restype = C.c_void_p
    L.ora_profile_new.argtypes = [C.c_size_t, C.c_size_t, C.c_int8]
    L.ora_profile_from_bytes.restype = C.c_void_p
    L.ora_profile_from_bytes.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t] + [C.c_int8] * 6
    L.ora_profile_free.argtypes = [C.c_void_p]
    L.ora_profile_len.restype = C.c_size_t
    L.ora_profile_len.argtypes = [C.c_void_p]
    L.ora_profile_clear.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    L.ora_profile_set.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_int8]
    L.ora_profile_set_all.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                      C.c_size_t, C.c_int]
    for nm in ("open_C", "close_C", "open_R"):
        getattr(L, f"ora_profile_set_gap_{nm}").argtypes = [C.c_void_p, C.c_size_t, C.c_int8]
        getattr(L, f"ora_profile_set_all_gap_{nm}").argtypes = [C.c_void_p, C.c_int8]
        getattr(L, f"ora_profile_gap_{nm}").restype = C.c_void_p
        getattr(L, f"ora_profile_gap_{nm}").argtypes = [C.c_void_p]
    L.ora_profile_get.restype = C.c_int8
    L.ora_profile_get.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8]
    L.ora_profile_get_gap_extend.restype = C.c_int8
    L.ora_profile_get_gap_extend.argtypes = [C.c_void_p]
    L.ora_profile_pos_aa.restype = C.c_void_p
    L.ora_profile_pos_aa.argtypes = [C.c_void_p]
    L.ora_profile_curr_len.restype = C.c_size_t
    L.ora_profile_curr_len.argtypes = [C.c_void_p]
    L.ora_nuc_matrix_simple.argtypes = [C.c_int8, C.c_int8, C.c_void_p]
    L.ora_aa_matrix_simple.argtypes = [C.c_int8, C.c_int8, C.c_void_p]
    L.ora_pad.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    L.ora_prefix_scan.argtypes = [C.c_void_p, C.c_int16, C.c_void_p]
    L.ora_prefix_scan_consts.argtypes = [C.c_int16, C.c_void_p, C.c_void_p]
    L.ora_batch_align.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ora_pad_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ora_last_batch_seconds.restype = C.c_double
    L.ora_hw_threads.restype = C.c_int
    _LIB = L
    return L


def nuc_matrix(match, mismatch):
    m = np.zeros(128, dtype=np.int8)
    lib().ora_nuc_matrix_simple(match, mismatch, m.ctypes.data)
    return m


def aa_matrix_simple(match, mismatch):
    m = np.zeros(27 * 32, dtype=np.int8)
    lib().ora_aa_matrix_simple(match, mismatch, m.ctypes.data)
    return m


def _load_builtin_matrices():
    """Parse block_aligner_b200/csrc/matrices_data.h (generated by tools/gen_matrices.py)."""
    import re
    txt = open(os.path.join(ROOT, "block_aligner_b200", "csrc", "matrices_data.h")).read()
    out = {}
    for m in re.finditer(r"#define BA_(\w+)_VALUES \\\n((?:.*\\\n)+)", txt):
        vals = [int(v) for v in re.findall(r"-?\d+", m.group(2))]
        assert len(vals) == 27 * 32
        out[m.group(1)] = np.array(vals, dtype=np.int8)
    return out


_BUILTIN = None


def builtin(name):
    global _BUILTIN
    if _BUILTIN is None:
        _BUILTIN = _load_builtin_matrices()
    return _BUILTIN[name]


NW1 = None


def nw1():
    global NW1
    if NW1 is None:
        NW1 = nuc_matrix(1, -1)
    return NW1


class Padded:
    """PaddedBytes::from_bytes::<M>(b, block_size) (scan_block.rs:1829-1836)."""

    def __init__(self, kind, b, block_size, rev=False):
        if isinstance(b, str):
            b = b.encode()
        self.kind, self.raw, self.n = kind, bytes(b), len(b)
        self.buf = np.zeros(1 + len(b) + block_size, dtype=np.uint8)
        e = lib().ora_pad(kind, self.raw, len(b), block_size, int(rev), self.buf.ctypes.data)
        if e:
            raise ValueError(f"ora_pad error {e}")

    def __len__(self):
        return self.n

    @property
    def ptr(self):
        return self.buf.ctypes.data


class Profile:
    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_bytes(cls, b, block_size, match, mismatch, gap_open_C, gap_close_C, gap_open_R, gap_extend):
        return cls(lib().ora_profile_from_bytes(b, len(b), block_size, match, mismatch, gap_open_C, gap_close_C,
                                                gap_open_R, gap_extend))

    @classmethod
    def new(cls, str_len, block_size, gap_extend):
        return cls(lib().ora_profile_new(str_len, block_size, gap_extend))

    def __len__(self):
        return lib().ora_profile_len(self.h)

    def set(self, i, b, score):
        assert lib().ora_profile_set(self.h, i, b, score) == 0

    def set_gap_open_C(self, i, g):
        assert lib().ora_profile_set_gap_open_C(self.h, i, g) == 0

    def set_gap_close_C(self, i, g):
        assert lib().ora_profile_set_gap_close_C(self.h, i, g) == 0

    def set_gap_open_R(self, i, g):
        assert lib().ora_profile_set_gap_open_R(self.h, i, g) == 0

    def __del__(self):
        try:
            lib().ora_profile_free(self.h)
        except Exception:
            pass


class Block:
    """Block::<TRACE, X_DROP, LOCAL_START, FREE_QUERY_START_GAPS, FREE_QUERY_END_GAPS>::new."""

    def __init__(self, query_len, reference_len, max_size, flags=0):
        self.flags = flags
        self.h = lib().ora_block_new(query_len, reference_len, max_size, flags)
        if not self.h:
            raise ValueError("Block size must be a power of two!")
        self._q = self._r = None

    def align(self, q, r, kind, matrix, gaps, size, x_drop=0):
        matrix = np.ascontiguousarray(matrix, dtype=np.int8)
        e = lib().ora_align(self.h, q.ptr, len(q), r.ptr, len(r), kind, matrix.ctypes.data, gaps[0], gaps[1],
                            size[0], size[1], x_drop)
        if e:
            raise ValueError(f"ora_align error {e}")
        self._q, self._r = q, r
        return self.res()

    def align_profile(self, q, prof, size, x_drop=0):
        e = lib().ora_align_profile(self.h, q.ptr, len(q), prof.h, size[0], size[1], x_drop)
        if e:
            raise ValueError(f"ora_align_profile error {e}")
        self._q, self._r = q, None
        self._rlen = len(prof)
        return self.res()

    def res(self):
        return lib().ora_res(self.h).tup()

    def cells(self):
        return lib().ora_cells(self.h)

    def steps(self):
        return lib().ora_steps(self.h)

    def cigar_runs(self, qi, rj, eq=False):
        rlen = len(self._r) if self._r is not None else self._rlen
        cap = len(self._q) + rlen + 5
        buf = (OpLen * cap)()
        n = lib().ora_cigar(self.h, qi, rj, int(eq), self._q.ptr, len(self._q),
                            self._r.ptr if self._r is not None else None, rlen, buf, cap)
        if n < 0:
            raise ValueError("cigar error")
        return [(buf[k].op, buf[k].len) for k in range(n)]

    def cigar(self, qi, rj, eq=False):
        return "".join(f"{ln}{OPS[op]}" for op, ln in self.cigar_runs(qi, rj, eq))

    def enable_step_log(self, on=True):
        lib().ora_enable_step_log(self.h, int(on))

    def logged_steps(self):
        n = lib().ora_align_logged_steps(self.h, None, 0)
        buf = (Step * max(n, 1))()
        lib().ora_align_logged_steps(self.h, buf, n)
        return [(s.dir, s.i, s.j, s.block_size, s.off, s.max, s.right_max, s.down_max) for s in buf[:n]]

    def __del__(self):
        try:
            lib().ora_block_free(self.h)
        except Exception:
            pass


def prefix_scan(vec16, gap):
    a = np.ascontiguousarray(vec16, dtype=np.int16)
    o = np.zeros(16, dtype=np.int16)
    lib().ora_prefix_scan(a.ctypes.data, gap, o.ctypes.data)
    return o
