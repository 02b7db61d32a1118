"""ctypes binding of libblock_aligner_b200.so plus a thin mirror of the reference's Rust API.

Names follow the reference (`PaddedBytes`, `NucMatrix`/`AAMatrix`/`AAProfile`, `Gaps`,
`Block(trace, x_drop).align`, `AlignResult`, `Cigar`; reference: src/scan_block.rs, src/scores.rs,
src/cigar.rs) so that tests read like the reference's own. All compute happens in the CUDA library;
this module only marshals. There is no CPU fallback: if the shared library is missing or no GPU is
usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libblock_aligner_b200.so")

SCORING_NUC, SCORING_AA, SCORING_BYTE, SCORING_PROFILE = 0, 1, 2, 3
TRACE, XDROP, LOCAL_START, FREE_QUERY_START_GAPS, FREE_QUERY_END_GAPS = 1, 2, 4, 8, 16
REV_QUERY, REV_REFERENCE = 32, 64      # PaddedBytes::set_bytes_rev for the whole batch (done on the device)
INPUT_NUC4 = 128                       # arenas hold BAM 4-bit codes, offsets count nibbles
OPS = {0: "?", 1: "M", 2: "=", 3: "X", 4: "I", 5: "D"}


class Gaps(C.Structure):
    """scores.rs:335-338"""
    _fields_ = [("open", C.c_int8), ("extend", C.c_int8)]


class SizeRange(C.Structure):
    """ffi.rs:20-23"""
    _fields_ = [("min", C.c_size_t), ("max", C.c_size_t)]


class AlignResult(C.Structure):
    """scan_block.rs:1887-1893"""
    _fields_ = [("score", C.c_int32), ("query_idx", C.c_size_t), ("reference_idx", C.c_size_t)]

    def tup(self):
        return (self.score, self.query_idx, self.reference_idx)


class OpLen(C.Structure):
    """cigar.rs:36-39"""
    _fields_ = [("op", C.c_uint8), ("len", C.c_size_t)]


class BaConfig(C.Structure):
    _fields_ = [("scoring", C.c_int32), ("flags", C.c_int32), ("matrix", C.c_void_p), ("gaps", Gaps),
                ("size", SizeRange), ("x_drop", C.c_int32), ("cigar_eq", C.c_int32)]


class BaStats(C.Structure):
    _fields_ = [("cells", C.c_uint64), ("steps", C.c_uint64), ("kernel_ms", C.c_float), ("pack_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("n_failed", C.c_uint32)]


class BaPssmBatch(C.Structure):
    _fields_ = [("order", C.c_void_p), ("order_len", C.c_size_t), ("scores", C.c_void_p), ("score_off", C.c_void_p),
                ("left_shift", C.c_size_t), ("right_shift", C.c_size_t), ("rev", C.c_int32),
                ("gap_open_C", C.c_void_p), ("gap_close_C", C.c_void_p), ("gap_open_R", C.c_void_p), ("gap_off", C.c_void_p),
                ("all_gap_open_C", C.c_int8), ("all_gap_close_C", C.c_int8), ("all_gap_open_R", C.c_int8),
                ("gap_extend", C.c_int8)]


class StepLog(C.Structure):
    _fields_ = [("dir", C.c_int32), ("i", C.c_uint32), ("j", C.c_uint32), ("block_size", C.c_uint32),
                ("off", C.c_int32), ("max", C.c_int16), ("right_max", C.c_int16), ("down_max", C.c_int16)]


class BlockAlignerError(RuntimeError):
    pass


_BLOCK_SUFFIX = {0: "aa", XDROP: "aa_xdrop", TRACE: "aa_trace", TRACE | XDROP: "aa_trace_xdrop"}
_CIGAR_FN = {0: "_block_cigar_aa", XDROP: "_block_cigar_aa_xdrop", TRACE: "block_cigar_aa_trace",
             TRACE | XDROP: "block_cigar_aa_trace_xdrop"}
_CIGAR_EQ_FN = {0: "_block_cigar_eq_aa", XDROP: "_block_cigar_eq_aa_xdrop", TRACE: "block_cigar_eq_aa_trace",
                TRACE | XDROP: "block_cigar_eq_aa_trace_xdrop"}

# every symbol include/block_aligner_b200.h declares (checked by tests/test_cabi_symbols.py)
PART1_FUNCTIONS = (
    ["block_new_simple_aamatrix", "block_set_aamatrix", "block_free_aamatrix", "block_new_aaprofile",
     "block_len_aaprofile", "block_clear_aaprofile", "block_set_aaprofile", "block_set_all_aaprofile",
     "block_set_all_rev_aaprofile", "block_set_gap_open_C_aaprofile", "block_set_gap_close_C_aaprofile",
     "block_set_gap_open_R_aaprofile", "block_set_all_gap_open_C_aaprofile", "block_set_all_gap_close_C_aaprofile",
     "block_set_all_gap_open_R_aaprofile", "block_get_aaprofile", "block_get_gap_extend_aaprofile",
     "block_free_aaprofile", "block_new_cigar", "block_get_cigar", "block_len_cigar", "block_free_cigar",
     "block_new_padded_aa", "block_set_bytes_padded_aa", "block_set_bytes_rev_padded_aa", "block_free_padded_aa"]
    + [f"block_{op}_{sfx}" for sfx in _BLOCK_SUFFIX.values() for op in ("new", "align", "align_profile", "res", "free")]
    + list(_CIGAR_FN.values()) + list(_CIGAR_EQ_FN.values()))
PART1_DATA = ["NW1", "BLOSUM45", "BLOSUM50", "BLOSUM62", "BLOSUM80", "BLOSUM90", "PAM100", "PAM120", "PAM160",
              "PAM200", "PAM250", "BYTES1"]
PART2_FUNCTIONS = ["ba_error_string", "ba_last_error_message", "ba_create", "ba_destroy", "ba_trim", "ba_batch_upload",
                   "ba_batch_upload_profiles", "ba_batch_upload_pssm", "ba_align_batch_pssm", "ba_batch_run", "ba_batch_download", "ba_batch_cigar",
                   "ba_batch_traceback", "ba_batch_total_stats", "ba_batch_pair_stats", "ba_batch_free",
                   "ba_align_batch", "ba_align_batch_cigar", "ba_align_batch_exp", "ba_align_batch_exp_profiles", "ba_align_batch_exp_pssm", "ba_align_batch_profiles", "ba_new_simple_nucmatrix", "ba_set_nucmatrix", "ba_free_nucmatrix",
                   "ba_device_count", "ba_multi_release", "ba_align_batch_multi", "ba_align_batch_multi_cigar", "ba_align_batch_multi_pssm",
                   "ba_percent_len", "ba_pack_nuc4", "ba_cigar_format", "ba_measure_int_peak", "ba_measure_int_peak_packed"]


class Library:
    """One loaded copy of the C-ABI library."""

    def __init__(self, path=None):
        path = path or os.environ.get("BLOCK_ALIGNER_B200_LIB") or DEFAULT_LIB
        if not os.path.exists(path):
            raise BlockAlignerError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        self.path = path
        L = self.L = C.CDLL(path)
        vp, sz, i8, u8, i32 = C.c_void_p, C.c_size_t, C.c_int8, C.c_uint8, C.c_int32
        L.ba_error_string.restype = C.c_char_p
        L.ba_error_string.argtypes = [C.c_int]
        L.ba_last_error_message.restype = C.c_char_p
        L.ba_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.ba_destroy.argtypes = [vp]
        L.ba_trim.argtypes = [vp]
        L.ba_batch_upload.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, vp, vp, C.POINTER(vp)]
        L.ba_batch_upload_profiles.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, vp, C.POINTER(vp)]
        L.ba_batch_run.argtypes = [vp, C.POINTER(BaStats)]
        L.ba_batch_download.argtypes = [vp, vp]
        L.ba_batch_cigar.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(sz)]
        L.ba_batch_traceback.argtypes = [vp, sz, sz, sz, C.c_int, C.POINTER(vp), C.POINTER(sz)]
        L.ba_batch_total_stats.argtypes = [vp, C.POINTER(BaStats)]
        L.ba_batch_pair_stats.argtypes = [vp, sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.ba_batch_free.argtypes = [vp]
        L.ba_align_batch.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, vp, vp, vp, C.POINTER(BaStats)]
        L.ba_align_batch_exp.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, vp, vp, vp, vp, vp, C.POINTER(BaStats)]
        L.ba_align_batch_profiles.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, vp, vp, C.POINTER(BaStats)]
        L.ba_align_batch_exp_profiles.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, vp, vp, vp, vp, C.POINTER(BaStats)]
        L.ba_align_batch_exp_pssm.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, C.POINTER(BaPssmBatch), vp, vp, vp, C.POINTER(BaStats)]
        L.ba_align_batch_cigar.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, vp, vp, vp, vp, sz, vp, vp, C.POINTER(sz),
                                           C.POINTER(BaStats)]
        L.ba_batch_upload_pssm.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, C.POINTER(BaPssmBatch), C.POINTER(vp)]
        L.ba_align_batch_pssm.argtypes = [vp, C.POINTER(BaConfig), sz, vp, vp, C.POINTER(BaPssmBatch), vp, C.POINTER(BaStats)]
        L.ba_new_simple_nucmatrix.restype = vp
        L.ba_new_simple_nucmatrix.argtypes = [i8, i8]
        L.ba_set_nucmatrix.argtypes = [vp, u8, u8, i8]
        L.ba_free_nucmatrix.argtypes = [vp]
        L.ba_device_count.restype = C.c_int
        L.ba_multi_release.restype = None
        L.ba_align_batch_multi.argtypes = [vp, C.c_int, C.POINTER(BaConfig), sz, vp, vp, vp, vp, vp, C.POINTER(BaStats)]
        L.ba_align_batch_multi_cigar.argtypes = [vp, C.c_int, C.POINTER(BaConfig), sz, vp, vp, vp, vp, vp, vp, sz, vp, vp, C.POINTER(sz),
                                                 C.POINTER(BaStats)]
        L.ba_align_batch_multi_pssm.argtypes = [vp, C.c_int, C.POINTER(BaConfig), sz, vp, vp, C.POINTER(BaPssmBatch), vp, C.POINTER(BaStats)]
        L.ba_pack_nuc4.restype = sz
        L.ba_pack_nuc4.argtypes = [vp, sz, vp, C.c_uint64]
        L.ba_percent_len.restype = sz
        L.ba_percent_len.argtypes = [sz, C.c_float]
        L.ba_cigar_format.restype = sz
        L.ba_cigar_format.argtypes = [vp, sz, C.c_char_p, sz]
        L.ba_measure_int_peak.argtypes = [vp, C.POINTER(C.c_double)]
        L.ba_measure_int_peak_packed.argtypes = [vp, C.POINTER(C.c_double)]
        if hasattr(L, "ba_debug_step_log"):
            L.ba_debug_step_log.restype = sz
            L.ba_debug_step_log.argtypes = [vp, vp, sz]
        # Part 1
        L.block_new_simple_aamatrix.restype = vp
        L.block_new_simple_aamatrix.argtypes = [i8, i8]
        L.block_set_aamatrix.argtypes = [vp, u8, u8, i8]
        L.block_free_aamatrix.argtypes = [vp]
        L.block_new_aaprofile.restype = vp
        L.block_new_aaprofile.argtypes = [sz, sz, i8]
        L.block_len_aaprofile.restype = sz
        L.block_len_aaprofile.argtypes = [vp]
        L.block_clear_aaprofile.argtypes = [vp, sz, sz]
        L.block_set_aaprofile.argtypes = [vp, sz, u8, i8]
        L.block_set_all_aaprofile.argtypes = [vp, vp, sz, vp, sz, sz, sz]
        L.block_set_all_rev_aaprofile.argtypes = [vp, vp, sz, vp, sz, sz, sz]
        for nm in ("open_C", "close_C", "open_R"):
            getattr(L, f"block_set_gap_{nm}_aaprofile").argtypes = [vp, sz, i8]
            getattr(L, f"block_set_all_gap_{nm}_aaprofile").argtypes = [vp, i8]
        L.block_get_aaprofile.restype = i8
        L.block_get_aaprofile.argtypes = [vp, sz, u8]
        L.block_get_gap_extend_aaprofile.restype = i8
        L.block_get_gap_extend_aaprofile.argtypes = [vp]
        L.block_free_aaprofile.argtypes = [vp]
        L.block_new_cigar.restype = vp
        L.block_new_cigar.argtypes = [sz, sz]
        L.block_get_cigar.restype = OpLen
        L.block_get_cigar.argtypes = [vp, sz]
        L.block_len_cigar.restype = sz
        L.block_len_cigar.argtypes = [vp]
        L.block_free_cigar.argtypes = [vp]
        L.block_new_padded_aa.restype = vp
        L.block_new_padded_aa.argtypes = [sz, sz]
        L.block_set_bytes_padded_aa.argtypes = [vp, C.c_char_p, sz, sz]
        L.block_set_bytes_rev_padded_aa.argtypes = [vp, C.c_char_p, sz, sz]
        L.block_free_padded_aa.argtypes = [vp]
        for fl, sfx in _BLOCK_SUFFIX.items():
            getattr(L, f"block_new_{sfx}").restype = vp
            getattr(L, f"block_new_{sfx}").argtypes = [sz, sz, sz]
            getattr(L, f"block_align_{sfx}").argtypes = [vp, vp, vp, vp, Gaps, SizeRange, i32]
            getattr(L, f"block_align_profile_{sfx}").argtypes = [vp, vp, vp, SizeRange, i32]
            getattr(L, f"block_res_{sfx}").restype = AlignResult
            getattr(L, f"block_res_{sfx}").argtypes = [vp]
            getattr(L, f"block_free_{sfx}").argtypes = [vp]
            getattr(L, _CIGAR_FN[fl]).argtypes = [vp, sz, sz, vp]
            getattr(L, _CIGAR_EQ_FN[fl]).argtypes = [vp, vp, vp, sz, sz, vp]

    def check(self, rc):
        if rc != 0:
            raise BlockAlignerError(f"{self.L.ba_error_string(rc).decode()} [{self.L.ba_last_error_message().decode()}]")

    def data_symbol(self, name, nbytes):
        return np.frombuffer((C.c_int8 * nbytes).in_dll(self.L, name), dtype=np.int8)

    def builtin_matrix(self, name):
        """NW1 -> (SCORING_NUC, 128 i8); BLOSUM*/PAM* -> (SCORING_AA, 864 i8); BYTES1 -> (SCORING_BYTE, 2 i8)."""
        if name == "NW1":
            return SCORING_NUC, self.data_symbol(name, 128)
        if name == "BYTES1":
            return SCORING_BYTE, self.data_symbol(name, 2)
        return SCORING_AA, self.data_symbol(name, 27 * 32)

    def percent_len(self, length, p):
        return self.L.ba_percent_len(length, p)


def nuc_matrix(match, mismatch):
    """NucMatrix::new_simple (scores.rs:150-164) as a 128-entry i8 array."""
    m = np.full(128, -128, dtype=np.int8)
    for i, a in enumerate(b"ATCGN"):
        for j, b in enumerate(b"ATCGN"):
            m[(a & 7) * 16 + (b & 15)] = match if i == j else mismatch
    return m


def aa_matrix_simple(match, mismatch):
    """AAMatrix::new_simple (scores.rs:48-61)."""
    m = np.full(27 * 32, -128, dtype=np.int8)
    for i in range(26):
        for j in range(26):
            m[i * 32 + j] = match if i == j else mismatch
    return m


def concat(seqs):
    """list of bytes -> (uint8 arena, uint64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    arena = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(arena), off


def pack_nuc4(lib, arena, off):
    """ASCII arena + byte offsets -> (BAM-nibble arena, the same offsets read as nibble offsets) for INPUT_NUC4:
    the concatenation stays dense, sequence k = nibbles [off[k], off[k+1])."""
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    total = int(off[-1])
    packed = np.zeros((total + 1) // 2 + 1, dtype=np.uint8)
    bad = lib.L.ba_pack_nuc4(arena.ctypes.data, total, packed.ctypes.data, 0)
    if bad:
        raise BlockAlignerError(f"byte {bad - 1} is not a nucleotide code")
    return packed, np.ascontiguousarray(off, dtype=np.uint64)


def runs_to_string(runs):
    return "".join(f"{int(r) >> 4}{OPS[int(r) & 15]}" for r in runs)


class Aligner:
    """BaAligner: one per GPU."""

    def __init__(self, lib=None, device=0):
        self.lib = lib or Library()
        h = C.c_void_p()
        self.lib.check(self.lib.L.ba_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.L.ba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def int_peak_gops(self, packed=False):
        v = C.c_double()
        fn = self.lib.L.ba_measure_int_peak_packed if packed else self.lib.L.ba_measure_int_peak
        self.lib.check(fn(self.h, C.byref(v)))
        return v.value

    def config(self, scoring, matrix, gaps, size, x_drop=0, flags=0, cigar_eq=False):
        cfg = BaConfig()
        cfg.scoring, cfg.flags = scoring, flags
        self._mat = None if matrix is None else np.ascontiguousarray(matrix, dtype=np.int8)
        cfg.matrix = None if matrix is None else self._mat.ctypes.data
        cfg.gaps = Gaps(*gaps) if gaps else Gaps(0, 0)
        cfg.size = SizeRange(*size)
        cfg.x_drop, cfg.cigar_eq = x_drop, int(bool(cigar_eq))
        return cfg

    def upload(self, cfg, q_arena, q_off, r_arena=None, r_off=None, profiles=None):
        return Batch(self, cfg, q_arena, q_off, r_arena, r_off, profiles)

    def align_batch(self, queries, references, scoring, matrix, gaps, size, x_drop=0, flags=0, cigar_eq=False):
        """Convenience: lists of bytes in, list of (score, query_idx, reference_idx) [+ CIGAR strings] out."""
        qa, qo = concat(queries)
        ra, ro = concat(references)
        cfg = self.config(scoring, matrix, gaps, size, x_drop, flags, cigar_eq)
        b = self.upload(cfg, qa, qo, ra, ro)
        try:
            b.run()
            res = b.download()
            out = [tuple(int(x) for x in r) for r in res]
            cig = [b.cigar_string(k) for k in range(len(queries))] if flags & TRACE else None
            return out, cig, b.total_stats()
        finally:
            b.free()


def align_batch_exp(al, queries, references, scoring, matrix, gaps, size, target_scores, x_drop=0, flags=0, profiles=None):
    """Block::align_exp / align_profile_exp for a batch -> (results, min_size_used with None where the target was never
    reached, BaStats). `profiles`: a PssmBatch or a list of AAProfile (then `references` is ignored)."""
    qa, qo = concat(queries)
    n = len(queries)
    cfg = al.config(scoring, matrix, gaps, size, x_drop, flags, False)
    out = np.zeros(max(n, 1), dtype=np.dtype([("score", np.int32), ("query_idx", np.uint64), ("reference_idx", np.uint64)], align=True))
    used = np.zeros(max(n, 1), dtype=np.uint64)
    tgt = np.ascontiguousarray(target_scores, dtype=np.int32)
    st = BaStats()
    L = al.lib.L
    if isinstance(profiles, PssmBatch):
        al.lib.check(L.ba_align_batch_exp_pssm(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, C.byref(profiles.c),
                                               tgt.ctypes.data, out.ctypes.data, used.ctypes.data, C.byref(st)))
    elif profiles is not None:
        arr = (C.c_void_p * n)(*[p.h for p in profiles])
        al.lib.check(L.ba_align_batch_exp_profiles(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, arr,
                                                   tgt.ctypes.data, out.ctypes.data, used.ctypes.data, C.byref(st)))
    else:
        ra, ro = concat(references)
        al.lib.check(L.ba_align_batch_exp(al.h, C.byref(cfg), n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data, ro.ctypes.data,
                                          tgt.ctypes.data, out.ctypes.data, used.ctypes.data, C.byref(st)))
    res = [(int(r["score"]), int(r["query_idx"]), int(r["reference_idx"])) for r in out[:n]]
    return res, [int(u) if u else None for u in used[:n]], st


class PssmBatch:
    """Raw PSSM rows for profiles built on the device (BaPssmBatch in include/block_aligner_b200.h): the batch form of
    AAProfile::new + set_all / set_all_rev + the gap setters (src/scores.rs:473-486, 532-538, 548-580).

    scores: int8 array, all profiles concatenated, position-major rows of len(order) entries; score_off: uint64[n + 1].
    gaps: None -> the three `all_gaps` values (open_C, close_C, open_R) apply to positions 0..=len; otherwise a tuple of
    three int8 arrays (open_C, close_C, open_R) with len + 1 entries per profile, concatenated."""

    def __init__(self, order, scores, score_off, gap_extend, all_gaps=(-10, 0, -10), gaps=None, left_shift=0, right_shift=0,
                 rev=False):
        self.order = np.frombuffer(bytes(order), dtype=np.uint8).copy()
        self.scores = scores if (isinstance(scores, np.ndarray) and scores.dtype == np.int8 and scores.flags.c_contiguous) \
            else np.ascontiguousarray(scores, dtype=np.int8)
        self.score_off = np.ascontiguousarray(score_off, dtype=np.uint64)
        self.n = len(self.score_off) - 1
        self.gaps = None
        self.gap_off = None
        c = BaPssmBatch()
        c.order, c.order_len = self.order.ctypes.data, len(self.order)
        c.scores, c.score_off = self.scores.ctypes.data, self.score_off.ctypes.data
        c.left_shift, c.right_shift, c.rev = left_shift, right_shift, int(bool(rev))
        if gaps is not None:
            self.gaps = tuple(np.ascontiguousarray(g, dtype=np.int8) for g in gaps)
            lens = (self.score_off[1:] - self.score_off[:-1]) // np.uint64(len(self.order))
            self.gap_off = np.zeros(self.n + 1, dtype=np.uint64)
            np.cumsum(lens + np.uint64(1), out=self.gap_off[1:])
            c.gap_open_C, c.gap_close_C, c.gap_open_R = (g.ctypes.data for g in self.gaps)
            c.gap_off = self.gap_off.ctypes.data
        c.all_gap_open_C, c.all_gap_close_C, c.all_gap_open_R = all_gaps
        c.gap_extend = gap_extend
        self.c = c

    def nbytes(self):
        return int(self.scores.nbytes + self.score_off.nbytes + (sum(g.nbytes for g in self.gaps) + self.gap_off.nbytes if self.gaps else 0))


class Batch:
    def __init__(self, al, cfg, q_arena, q_off, r_arena, r_off, profiles):
        self.al, self.cfg = al, cfg
        self.n = len(q_off) - 1
        self._keep = (q_arena, q_off, r_arena, r_off, profiles)
        h = C.c_void_p()
        L = al.lib.L
        qa = np.ascontiguousarray(q_arena, dtype=np.uint8)
        qo = np.ascontiguousarray(q_off, dtype=np.uint64)
        if isinstance(profiles, PssmBatch):
            al.lib.check(L.ba_batch_upload_pssm(al.h, C.byref(cfg), self.n, qa.ctypes.data, qo.ctypes.data, C.byref(profiles.c),
                                                C.byref(h)))
        elif profiles is not None:
            arr = (C.c_void_p * self.n)(*[p.h for p in profiles])
            al.lib.check(L.ba_batch_upload_profiles(al.h, C.byref(cfg), self.n, qa.ctypes.data, qo.ctypes.data, arr, C.byref(h)))
        else:
            ra = np.ascontiguousarray(r_arena, dtype=np.uint8)
            ro = np.ascontiguousarray(r_off, dtype=np.uint64)
            al.lib.check(L.ba_batch_upload(al.h, C.byref(cfg), self.n, qa.ctypes.data, qo.ctypes.data, ra.ctypes.data,
                                           ro.ctypes.data, C.byref(h)))
        self.h = h

    def run(self):
        st = BaStats()
        self.al.lib.check(self.al.lib.L.ba_batch_run(self.h, C.byref(st)))
        return st

    def download(self):
        """-> structured array with fields score, query_idx, reference_idx"""
        out = np.zeros(self.n, dtype=np.dtype([("score", np.int32), ("query_idx", np.uint64), ("reference_idx", np.uint64)],
                                              align=True))
        assert out.dtype.itemsize == C.sizeof(AlignResult)
        self.al.lib.check(self.al.lib.L.ba_batch_download(self.h, out.ctypes.data))
        return out

    def cigar_runs(self, k):
        p, n = C.c_void_p(), C.c_size_t()
        self.al.lib.check(self.al.lib.L.ba_batch_cigar(self.h, k, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=np.uint32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n.value,)).copy()

    def cigar_string(self, k):
        return runs_to_string(self.cigar_runs(k))

    def traceback(self, k, query_idx, reference_idx, eq=False):
        p, n = C.c_void_p(), C.c_size_t()
        self.al.lib.check(self.al.lib.L.ba_batch_traceback(self.h, k, query_idx, reference_idx, int(eq), C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=np.uint32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n.value,)).copy()

    def total_stats(self):
        st = BaStats()
        self.al.lib.check(self.al.lib.L.ba_batch_total_stats(self.h, C.byref(st)))
        return st

    def pair_stats(self, k):
        c, s, t = C.c_uint64(), C.c_uint32(), C.c_uint32()
        self.al.lib.check(self.al.lib.L.ba_batch_pair_stats(self.h, k, C.byref(c), C.byref(s), C.byref(t)))
        return c.value, s.value, t.value

    def step_log(self):
        cap = 1 << 20
        buf = (StepLog * cap)()
        n = self.al.lib.L.ba_debug_step_log(self.h, buf, cap)
        return [(s.dir, s.i, s.j, s.block_size, s.off, s.max, s.right_max, s.down_max) for s in buf[:min(n, cap)]]

    def free(self):
        if self.h:
            self.al.lib.L.ba_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------
# Mirror of the reference's per-pair API on top of Part 1 of the C ABI (what c/example.c uses)
# ---------------------------------------------------------------------------------------------------
class PaddedBytes:
    """PaddedBytes::from_bytes::<AAMatrix> (scan_block.rs:1829-1836) via block_new_padded_aa."""

    def __init__(self, lib, b, block_size):
        if isinstance(b, str):
            b = b.encode()
        self.lib, self.raw = lib, bytes(b)
        self.h = lib.L.block_new_padded_aa(len(b), block_size)
        lib.L.block_set_bytes_padded_aa(self.h, self.raw, len(b), block_size)

    def __len__(self):
        return len(self.raw)

    def __del__(self):
        try:
            self.lib.L.block_free_padded_aa(self.h)
        except Exception:
            pass


class AAProfile:
    """AAProfile (scores.rs:454-715) via the block_*_aaprofile calls."""

    def __init__(self, lib, str_len, block_size, gap_extend):
        self.lib = lib
        self.h = lib.L.block_new_aaprofile(str_len, block_size, gap_extend)

    @classmethod
    def from_bytes(cls, lib, b, block_size, match, mismatch, gap_open_C, gap_close_C, gap_open_R, gap_extend):
        """AAProfile::from_bytes (scores.rs:489-505)"""
        p = cls(lib, len(b), block_size, gap_extend)
        for i in range(len(b)):
            for c in range(ord("A"), ord("Z") + 1):
                lib.L.block_set_aaprofile(p.h, i + 1, c, match if c == b[i] else mismatch)
        for i in range(len(b) + 1):
            lib.L.block_set_gap_open_C_aaprofile(p.h, i, gap_open_C)
            lib.L.block_set_gap_close_C_aaprofile(p.h, i, gap_close_C)
            lib.L.block_set_gap_open_R_aaprofile(p.h, i, gap_open_R)
        return p

    def __len__(self):
        return self.lib.L.block_len_aaprofile(self.h)

    def set(self, i, b, score):
        self.lib.L.block_set_aaprofile(self.h, i, b, score)

    def set_all(self, order, scores, left_shift=0, right_shift=0, rev=False):
        sc = np.ascontiguousarray(scores, dtype=np.int8)
        od = np.frombuffer(bytes(order), dtype=np.uint8)
        fn = self.lib.L.block_set_all_rev_aaprofile if rev else self.lib.L.block_set_all_aaprofile
        fn(self.h, od.ctypes.data, len(od), sc.ctypes.data, sc.size, left_shift, right_shift)

    def set_gap_open_C(self, i, g):
        self.lib.L.block_set_gap_open_C_aaprofile(self.h, i, g)

    def set_gap_close_C(self, i, g):
        self.lib.L.block_set_gap_close_C_aaprofile(self.h, i, g)

    def set_gap_open_R(self, i, g):
        self.lib.L.block_set_gap_open_R_aaprofile(self.h, i, g)

    def set_all_gap_open_C(self, g):
        self.lib.L.block_set_all_gap_open_C_aaprofile(self.h, g)

    def set_all_gap_close_C(self, g):
        self.lib.L.block_set_all_gap_close_C_aaprofile(self.h, g)

    def set_all_gap_open_R(self, g):
        self.lib.L.block_set_all_gap_open_R_aaprofile(self.h, g)

    def get(self, i, b):
        return self.lib.L.block_get_aaprofile(self.h, i, b)

    def __del__(self):
        try:
            self.lib.L.block_free_aaprofile(self.h)
        except Exception:
            pass


class Cigar:
    def __init__(self, lib, query_len, reference_len):
        self.lib = lib
        self.h = lib.L.block_new_cigar(query_len, reference_len)

    def __len__(self):
        return self.lib.L.block_len_cigar(self.h)

    def to_vec(self):
        out = []
        for i in range(len(self)):
            o = self.lib.L.block_get_cigar(self.h, i)
            out.append((o.op, o.len))
        return out

    def to_string(self):
        return "".join(f"{n}{OPS[op]}" for op, n in self.to_vec())

    def __del__(self):
        try:
            self.lib.L.block_free_cigar(self.h)
        except Exception:
            pass


class Block:
    """Block::<TRACE, X_DROP> over amino-acid strings (the four monomorphs of ffi.rs:333-403)."""

    def __init__(self, lib, query_len, reference_len, max_size, trace=False, x_drop=False):
        self.lib = lib
        self.flags = (TRACE if trace else 0) | (XDROP if x_drop else 0)
        self.sfx = _BLOCK_SUFFIX[self.flags]
        self.h = getattr(lib.L, f"block_new_{self.sfx}")(query_len, reference_len, max_size)

    def align(self, q, r, matrix_ptr, gaps, size, x_drop=0):
        getattr(self.lib.L, f"block_align_{self.sfx}")(self.h, q.h, r.h, matrix_ptr, Gaps(*gaps), SizeRange(*size), x_drop)
        return self.res()

    def align_profile(self, q, prof, size, x_drop=0):
        getattr(self.lib.L, f"block_align_profile_{self.sfx}")(self.h, q.h, prof.h, SizeRange(*size), x_drop)
        return self.res()

    def res(self):
        return getattr(self.lib.L, f"block_res_{self.sfx}")(self.h).tup()

    def cigar(self, query_idx, reference_idx, cigar):
        getattr(self.lib.L, _CIGAR_FN[self.flags])(self.h, query_idx, reference_idx, cigar.h)
        return cigar.to_string()

    def cigar_eq(self, q, r, query_idx, reference_idx, cigar):
        getattr(self.lib.L, _CIGAR_EQ_FN[self.flags])(self.h, q.h, r.h, query_idx, reference_idx, cigar.h)
        return cigar.to_string()

    def __del__(self):
        try:
            getattr(self.lib.L, f"block_free_{self.sfx}")(self.h)
        except Exception:
            pass
