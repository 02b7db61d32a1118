//! Rust surface of the reference crate (`block_aligner::{scan_block, scores, cigar}`) over the C ABI of
//! `libblock_aligner_b200.so` (`include/block_aligner_b200.h`). Same type and method names as the reference:
//! `PaddedBytes::from_bytes::<M>`, `NucMatrix` / `AAMatrix` / `ByteMatrix`, `Gaps`,
//! `Block::<TRACE, X_DROP>::new / align / res / trace().cigar_eq`, `AlignResult`, `Cigar`, plus `align_batch`.
//!
//! NOT COMPILED in this repository's image (no Rust toolchain). The C++ header `include/block_aligner_b200.hpp`
//! is the compiled and tested twin of this file; keep the two in step.
#![allow(non_snake_case, clippy::too_many_arguments)]

use std::ffi::CStr;
use std::marker::PhantomData;
use std::ops::RangeInclusive;
use std::os::raw::{c_char, c_int, c_void};

// ---- C ABI (include/block_aligner_b200.h, Part 2) -------------------------------------------------------------
#[repr(C)]
#[derive(Clone, Copy, PartialEq, Debug)]
pub struct Gaps { pub open: i8, pub extend: i8 }                       // scores.rs:335-338
#[repr(C)]
#[derive(Clone, Copy)]
pub struct SizeRange { pub min: usize, pub max: usize }                // ffi.rs:20-23
#[repr(C)]
#[derive(Clone, Copy, PartialEq, Eq, Debug, Default)]
pub struct AlignResult { pub score: i32, pub query_idx: usize, pub reference_idx: usize }   // scan_block.rs:1887-1893

#[repr(C)]
struct BaConfig {
    scoring: i32, flags: i32, matrix: *const c_void, gaps: Gaps, size: SizeRange, x_drop: i32, cigar_eq: i32,
}
#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct BaStats { pub cells: u64, pub steps: u64, pub kernel_ms: f32, pub pack_ms: f32, pub kernel_launches: u32, pub n_failed: u32 }
enum BaAligner {}
enum BaBatch {}

const BA_TRACE: i32 = 1;
const BA_XDROP: i32 = 2;
const BA_LOCAL_START: i32 = 4;
const BA_FREE_QUERY_START_GAPS: i32 = 8;
const BA_FREE_QUERY_END_GAPS: i32 = 16;

extern "C" {
    fn ba_create(device: c_int, out: *mut *mut BaAligner) -> c_int;
    fn ba_destroy(a: *mut BaAligner);
    fn ba_error_string(code: c_int) -> *const c_char;
    fn ba_last_error_message() -> *const c_char;
    fn ba_batch_upload(a: *mut BaAligner, cfg: *const BaConfig, n: usize, q_bytes: *const u8, q_off: *const u64,
                       r_bytes: *const u8, r_off: *const u64, out: *mut *mut BaBatch) -> c_int;
    fn ba_batch_run(b: *mut BaBatch, stats: *mut BaStats) -> c_int;
    fn ba_batch_download(b: *mut BaBatch, out: *mut AlignResult) -> c_int;
    fn ba_batch_traceback(b: *mut BaBatch, k: usize, query_idx: usize, reference_idx: usize, eq: c_int,
                          runs: *mut *const u32, n_runs: *mut usize) -> c_int;
    fn ba_batch_free(b: *mut BaBatch);
    fn ba_align_batch(a: *mut BaAligner, cfg: *const BaConfig, n: usize, q_bytes: *const u8, q_off: *const u64,
                      r_bytes: *const u8, r_off: *const u64, out: *mut AlignResult, stats: *mut BaStats) -> c_int;
    fn ba_align_batch_cigar(a: *mut BaAligner, cfg: *const BaConfig, n: usize, q_bytes: *const u8, q_off: *const u64,
                            r_bytes: *const u8, r_off: *const u64, out: *mut AlignResult, runs: *mut u32, runs_cap: usize,
                            run_off: *mut u64, run_len: *mut u32, runs_used: *mut usize, stats: *mut BaStats) -> c_int;
    fn ba_align_batch_exp(a: *mut BaAligner, cfg: *const BaConfig, n: usize, q_bytes: *const u8, q_off: *const u64,
                          r_bytes: *const u8, r_off: *const u64, target_score: *const i32, out: *mut AlignResult,
                          min_size_used: *mut usize, stats: *mut BaStats) -> c_int;
    fn ba_device_count() -> c_int;
    fn ba_align_batch_multi(devices: *const c_int, n_dev: c_int, cfg: *const BaConfig, n: usize, q_bytes: *const u8,
                            q_off: *const u64, r_bytes: *const u8, r_off: *const u64, out: *mut AlignResult,
                            stats: *mut BaStats) -> c_int;
    fn ba_pack_nuc4(ascii: *const u8, len: usize, packed: *mut u8, nibble_off: u64) -> usize;
    fn ba_percent_len(len: usize, p: f32) -> usize;
    static NW1: NucMatrix;
    static BLOSUM62: AAMatrix;
    static BYTES1: ByteMatrix;
}

fn check(rc: c_int) {
    if rc != 0 {
        // the reference panics (and aborts in release) on every precondition violation; so does this binding
        let (a, b) = unsafe { (CStr::from_ptr(ba_error_string(rc)), CStr::from_ptr(ba_last_error_message())) };
        panic!("{}: {}", a.to_string_lossy(), b.to_string_lossy());
    }
}

/// One GPU. `Block`s share the process-wide default device; batches can name one.
pub struct Device { h: *mut BaAligner }
impl Device {
    pub fn new(index: i32) -> Self { let mut h = std::ptr::null_mut(); check(unsafe { ba_create(index, &mut h) }); Self { h } }
}
impl Drop for Device { fn drop(&mut self) { unsafe { ba_destroy(self.h) } } }
fn default_device() -> *mut BaAligner {
    use std::sync::OnceLock;
    struct P(*mut BaAligner);
    unsafe impl Send for P {}
    unsafe impl Sync for P {}
    static D: OnceLock<P> = OnceLock::new();
    D.get_or_init(|| { let mut h = std::ptr::null_mut(); check(unsafe { ba_create(0, &mut h) }); P(h) }).0
}

// ---- scores.rs --------------------------------------------------------------------------------------------------
pub trait Matrix {
    const NULL: u8;
    const SCORING: i32;
    fn as_ptr(&self) -> *const c_void;
    fn convert_char(c: u8) -> u8;
}
#[repr(C, align(32))]
#[derive(Clone, PartialEq, Debug)]
pub struct AAMatrix { scores: [i8; 27 * 32] }
#[repr(C, align(32))]
#[derive(Clone, PartialEq, Debug)]
pub struct NucMatrix { scores: [i8; 8 * 16] }
#[repr(C)]
#[derive(Clone, PartialEq, Debug)]
pub struct ByteMatrix { match_score: i8, mismatch_score: i8 }

impl AAMatrix {
    pub const fn new() -> Self { Self { scores: [i8::MIN; 27 * 32] } }
    pub fn new_simple(match_score: i8, mismatch_score: i8) -> Self {
        let mut m = Self::new();
        for a in 0..26 { for b in 0..26 { m.scores[a * 32 + b] = if a == b { match_score } else { mismatch_score }; } }
        m
    }
    pub fn set(&mut self, a: u8, b: u8, score: i8) {
        let (x, y) = (Self::convert_char(a) as usize, Self::convert_char(b) as usize);
        self.scores[x * 32 + y] = score; self.scores[y * 32 + x] = score;
    }
    pub fn get(&self, a: u8, b: u8) -> i8 { self.scores[Self::convert_char(a) as usize * 32 + Self::convert_char(b) as usize] }
}
impl Matrix for AAMatrix {
    const NULL: u8 = b'A' + 26;
    const SCORING: i32 = 1;
    fn as_ptr(&self) -> *const c_void { self.scores.as_ptr() as *const c_void }
    fn convert_char(c: u8) -> u8 { let c = c.to_ascii_uppercase(); assert!(c >= b'A' && c <= Self::NULL); c - b'A' }
}
impl NucMatrix {
    pub const fn new() -> Self { Self { scores: [i8::MIN; 8 * 16] } }
    pub fn new_simple(match_score: i8, mismatch_score: i8) -> Self {
        let mut m = Self::new();
        for &a in b"ACGNT" { for &b in b"ACGNT" { m.set(a, b, if a == b { match_score } else { mismatch_score }); } }
        m
    }
    pub fn set(&mut self, a: u8, b: u8, score: i8) {
        let (x, y) = (Self::convert_char(a) as usize, Self::convert_char(b) as usize);
        self.scores[(x & 7) * 16 + (y & 15)] = score; self.scores[(y & 7) * 16 + (x & 15)] = score;
    }
}
impl Matrix for NucMatrix {
    const NULL: u8 = b'Z';
    const SCORING: i32 = 0;
    fn as_ptr(&self) -> *const c_void { self.scores.as_ptr() as *const c_void }
    fn convert_char(c: u8) -> u8 { let c = c.to_ascii_uppercase(); assert!(c >= b'A' && c <= b'Z'); c }
}
impl ByteMatrix { pub const fn new_simple(match_score: i8, mismatch_score: i8) -> Self { Self { match_score, mismatch_score } } }
impl Matrix for ByteMatrix {
    const NULL: u8 = 0;
    const SCORING: i32 = 2;
    fn as_ptr(&self) -> *const c_void { self as *const Self as *const c_void }
    fn convert_char(c: u8) -> u8 { c }
}
pub fn nw1() -> &'static NucMatrix { unsafe { &NW1 } }
pub fn blosum62() -> &'static AAMatrix { unsafe { &BLOSUM62 } }
pub fn bytes1() -> &'static ByteMatrix { unsafe { &BYTES1 } }

/// lib.rs:109-111
pub fn percent_len(len: usize, p: f32) -> usize { unsafe { ba_percent_len(len, p) } }

// ---- scan_block.rs: PaddedBytes ---------------------------------------------------------------------------------
#[derive(Clone, PartialEq, Debug)]
pub struct PaddedBytes { s: Vec<u8>, raw: Vec<u8>, len: usize }
impl PaddedBytes {
    pub fn new<M: Matrix>(len: usize, block_size: usize) -> Self {
        Self { s: vec![M::convert_char(M::NULL); 1 + len + block_size], raw: Vec::with_capacity(len), len }
    }
    pub fn from_bytes<M: Matrix>(b: &[u8], block_size: usize) -> Self { let mut p = Self::new::<M>(b.len(), block_size); p.set_bytes::<M>(b, block_size); p }
    pub fn from_str<M: Matrix>(s: &str, block_size: usize) -> Self { Self::from_bytes::<M>(s.as_bytes(), block_size) }
    pub fn set_bytes<M: Matrix>(&mut self, b: &[u8], block_size: usize) { self.fill::<M>(b.iter().copied(), b.len(), block_size) }
    pub fn set_bytes_rev<M: Matrix>(&mut self, b: &[u8], block_size: usize) { self.fill::<M>(b.iter().rev().copied(), b.len(), block_size) }
    fn fill<M: Matrix>(&mut self, it: impl Iterator<Item = u8>, n: usize, block_size: usize) {
        assert!(1 + n + block_size <= self.s.len());
        let nul = M::convert_char(M::NULL);
        self.raw.clear();
        self.s[0] = nul;
        for (k, c) in it.enumerate() { self.raw.push(c); self.s[1 + k] = M::convert_char(c); }
        for k in 0..block_size { self.s[1 + n + k] = nul; }
        self.len = n;
    }
    pub fn get(&self, i: usize) -> u8 { self.s[i] }
    pub fn len(&self) -> usize { self.len }
}

// ---- cigar.rs ---------------------------------------------------------------------------------------------------
#[repr(u8)]
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum Operation { Sentinel = 0, M = 1, Eq = 2, X = 3, I = 4, D = 5 }
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub struct OpLen { pub op: Operation, pub len: usize }
#[derive(Clone, Debug, Default)]
pub struct Cigar { s: Vec<OpLen> }
impl Cigar {
    pub fn new(query_len: usize, reference_len: usize) -> Self { Self { s: Vec::with_capacity(query_len + reference_len + 5) } }
    pub fn len(&self) -> usize { self.s.len() }
    pub fn get(&self, i: usize) -> OpLen { self.s[i] }
    pub fn to_vec(&self) -> Vec<OpLen> { self.s.clone() }
    fn assign_runs(&mut self, runs: &[u32]) {
        self.s.clear();
        for &w in runs {
            let op = match w & 15 { 1 => Operation::M, 2 => Operation::Eq, 3 => Operation::X, 4 => Operation::I, 5 => Operation::D, _ => Operation::Sentinel };
            self.s.push(OpLen { op, len: (w >> 4) as usize });
        }
    }
}
impl std::fmt::Display for Cigar {     // cigar.rs:147-163
    fn fmt(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        for o in &self.s {
            let c = match o.op { Operation::M => 'M', Operation::Eq => '=', Operation::X => 'X', Operation::I => 'I', Operation::D => 'D', Operation::Sentinel => continue };
            write!(f, "{}{}", o.len, c)?;
        }
        Ok(())
    }
}

// ---- scan_block.rs: Block ---------------------------------------------------------------------------------------
pub struct Block<const TRACE: bool, const X_DROP: bool, const LOCAL_START: bool = false,
                 const FREE_QUERY_START_GAPS: bool = false, const FREE_QUERY_END_GAPS: bool = false> {
    query_len: usize, reference_len: usize, max_size: usize,
    res: AlignResult, last: *mut BaBatch, last_qlen: usize, last_rlen: usize,
}
pub struct Trace<'a> { b: *mut BaBatch, qlen: usize, rlen: usize, _p: PhantomData<&'a ()> }

impl<const TRACE: bool, const X_DROP: bool, const LOCAL_START: bool, const FREE_QUERY_START_GAPS: bool, const FREE_QUERY_END_GAPS: bool>
    Block<TRACE, X_DROP, LOCAL_START, FREE_QUERY_START_GAPS, FREE_QUERY_END_GAPS> {
    const FLAGS: i32 = (if TRACE { BA_TRACE } else { 0 }) | (if X_DROP { BA_XDROP } else { 0 }) | (if LOCAL_START { BA_LOCAL_START } else { 0 })
        | (if FREE_QUERY_START_GAPS { BA_FREE_QUERY_START_GAPS } else { 0 }) | (if FREE_QUERY_END_GAPS { BA_FREE_QUERY_END_GAPS } else { 0 });

    /// scan_block.rs:798-805
    pub fn new(query_len: usize, reference_len: usize, max_size: usize) -> Self {
        assert!(max_size.is_power_of_two(), "Block size must be a power of two!");
        Self { query_len, reference_len, max_size, res: AlignResult::default(), last: std::ptr::null_mut(), last_qlen: 0, last_rlen: 0 }
    }
    /// scan_block.rs:847-878
    pub fn align<M: Matrix>(&mut self, query: &PaddedBytes, reference: &PaddedBytes, matrix: &M, gaps: Gaps, size: RangeInclusive<usize>, x_drop: i32) {
        assert!(query.len() + reference.len() <= self.query_len + self.reference_len);
        assert!((*size.end()).max(16) <= self.max_size);
        let cfg = BaConfig { scoring: M::SCORING, flags: Self::FLAGS, matrix: matrix.as_ptr(), gaps,
                             size: SizeRange { min: *size.start(), max: *size.end() }, x_drop, cigar_eq: 0 };
        let (qo, ro) = ([0u64, query.len() as u64], [0u64, reference.len() as u64]);
        unsafe {
            if !self.last.is_null() { ba_batch_free(self.last); self.last = std::ptr::null_mut(); }
            check(ba_batch_upload(default_device(), &cfg, 1, query.raw.as_ptr(), qo.as_ptr(), reference.raw.as_ptr(), ro.as_ptr(), &mut self.last));
            check(ba_batch_run(self.last, std::ptr::null_mut()));
            check(ba_batch_download(self.last, &mut self.res));
        }
        self.last_qlen = query.len(); self.last_rlen = reference.len();
    }
    /// scan_block.rs:884-902
    pub fn align_exp<M: Matrix>(&mut self, query: &PaddedBytes, reference: &PaddedBytes, matrix: &M, gaps: Gaps, size: RangeInclusive<usize>,
                                x_drop: i32, target_score: i32) -> Option<usize> {
        let mut s = (*size.start()).max(16);   // scan_block.rs:885: `cmp::max(*size.start(), L)`, L = 16 on AVX2
        while s <= *size.end() {
            self.align(query, reference, matrix, gaps, s..=*size.end(), x_drop);
            if self.res.score >= target_score { return Some(s); }
            s *= 2;
        }
        None
    }
    pub fn res(&self) -> AlignResult { self.res }
    /// scan_block.rs:1241
    pub fn trace(&self) -> Trace<'_> { assert!(TRACE); Trace { b: self.last, qlen: self.last_qlen, rlen: self.last_rlen, _p: PhantomData } }
}
impl<const A: bool, const B: bool, const C: bool, const D: bool, const E: bool> Drop for Block<A, B, C, D, E> {
    fn drop(&mut self) { if !self.last.is_null() { unsafe { ba_batch_free(self.last) } } }
}
impl Trace<'_> {
    fn walk(&self, query_idx: usize, reference_idx: usize, eq: bool, cigar: &mut Cigar) {
        assert!(query_idx <= self.qlen && reference_idx <= self.rlen, "Traceback cigar end position must be in bounds!");
        let (mut runs, mut n) = (std::ptr::null(), 0usize);
        check(unsafe { ba_batch_traceback(self.b, 0, query_idx, reference_idx, eq as c_int, &mut runs, &mut n) });
        cigar.assign_runs(unsafe { std::slice::from_raw_parts(runs, n) });
    }
    /// scan_block.rs:1469-1473
    pub fn cigar(&self, query_idx: usize, reference_idx: usize, cigar: &mut Cigar) { self.walk(query_idx, reference_idx, false, cigar) }
    /// scan_block.rs:1475-1480
    pub fn cigar_eq(&self, _q: &PaddedBytes, _r: &PaddedBytes, query_idx: usize, reference_idx: usize, cigar: &mut Cigar) {
        self.walk(query_idx, reference_idx, true, cigar)
    }
}

// ---- batch extension ----------------------------------------------------------------------------------------------
/// `for (q, r) in pairs { block.align(q, r, ..) }` as one call; with TRACE the CIGAR strings come back too.
pub fn align_batch<const TRACE: bool, const X_DROP: bool, M: Matrix>(
    dev: &Device, queries: &[&[u8]], references: &[&[u8]], matrix: &M, gaps: Gaps, size: RangeInclusive<usize>, x_drop: i32,
) -> (Vec<AlignResult>, Vec<String>, BaStats) {
    assert_eq!(queries.len(), references.len());
    let n = queries.len();
    let offsets = |v: &[&[u8]]| { let mut o = vec![0u64; n + 1]; for k in 0..n { o[k + 1] = o[k] + v[k].len() as u64; } o };
    let (qo, ro) = (offsets(queries), offsets(references));
    let (qa, ra): (Vec<u8>, Vec<u8>) = (queries.concat(), references.concat());
    let cfg = BaConfig { scoring: M::SCORING, flags: (if TRACE { BA_TRACE } else { 0 }) | (if X_DROP { BA_XDROP } else { 0 }),
                         matrix: matrix.as_ptr(), gaps, size: SizeRange { min: *size.start(), max: *size.end() }, x_drop, cigar_eq: 1 };
    let mut out = vec![AlignResult::default(); n];
    let mut stats = BaStats::default();
    let mut cigars = Vec::new();
    if TRACE {
        let cap = (qo[n] + ro[n]) as usize + 5 * n;
        let (mut runs, mut off, mut len, mut used) = (vec![0u32; cap], vec![0u64; n], vec![0u32; n], 0usize);
        check(unsafe { ba_align_batch_cigar(dev.h, &cfg, n, qa.as_ptr(), qo.as_ptr(), ra.as_ptr(), ro.as_ptr(), out.as_mut_ptr(),
                                            runs.as_mut_ptr(), cap, off.as_mut_ptr(), len.as_mut_ptr(), &mut used, &mut stats) });
        for k in 0..n {
            let mut c = Cigar::default();
            c.assign_runs(&runs[off[k] as usize..off[k] as usize + len[k] as usize]);
            cigars.push(c.to_string());
        }
    } else {
        check(unsafe { ba_align_batch(dev.h, &cfg, n, qa.as_ptr(), qo.as_ptr(), ra.as_ptr(), ro.as_ptr(), out.as_mut_ptr(), &mut stats) });
    }
    (out, cigars, stats)
}

/// One batch over several GPUs of the box in one call (`devices` empty: every visible device). Contiguous shards of equal
/// sum(|q| + |r|), one host thread per GPU inside the library, results in the caller's order; no collective.
pub fn align_batch_multi<const X_DROP: bool, M: Matrix>(
    devices: &[i32], queries: &[&[u8]], references: &[&[u8]], matrix: &M, gaps: Gaps, size: RangeInclusive<usize>, x_drop: i32,
) -> (Vec<AlignResult>, BaStats) {
    assert_eq!(queries.len(), references.len());
    let n = queries.len();
    let offsets = |v: &[&[u8]]| { let mut o = vec![0u64; n + 1]; for k in 0..n { o[k + 1] = o[k] + v[k].len() as u64; } o };
    let (qo, ro) = (offsets(queries), offsets(references));
    let (qa, ra): (Vec<u8>, Vec<u8>) = (queries.concat(), references.concat());
    let cfg = BaConfig { scoring: M::SCORING, flags: if X_DROP { BA_XDROP } else { 0 }, matrix: matrix.as_ptr(), gaps,
                         size: SizeRange { min: *size.start(), max: *size.end() }, x_drop, cigar_eq: 0 };
    let mut out = vec![AlignResult::default(); n];
    let mut stats = BaStats::default();
    let (dp, nd) = if devices.is_empty() { (std::ptr::null(), unsafe { ba_device_count() }) } else { (devices.as_ptr(), devices.len() as c_int) };
    check(unsafe { ba_align_batch_multi(dp, nd, &cfg, n, qa.as_ptr(), qo.as_ptr(), ra.as_ptr(), ro.as_ptr(), out.as_mut_ptr(), &mut stats) });
    (out, stats)
}

/// `Block::align_exp` (scan_block.rs:884-902) for a batch: `Some(min size that reached target_scores[k])` or `None` per pair.
/// The batch stays on the device between the rounds; only the pairs still below their target are aligned again.
pub fn align_batch_exp<const X_DROP: bool, M: Matrix>(
    dev: &Device, queries: &[&[u8]], references: &[&[u8]], matrix: &M, gaps: Gaps, size: RangeInclusive<usize>, x_drop: i32,
    target_scores: &[i32],
) -> (Vec<AlignResult>, Vec<Option<usize>>) {
    assert!(queries.len() == references.len() && queries.len() == target_scores.len());
    let n = queries.len();
    let offsets = |v: &[&[u8]]| { let mut o = vec![0u64; n + 1]; for k in 0..n { o[k + 1] = o[k] + v[k].len() as u64; } o };
    let (qo, ro) = (offsets(queries), offsets(references));
    let (qa, ra): (Vec<u8>, Vec<u8>) = (queries.concat(), references.concat());
    let cfg = BaConfig { scoring: M::SCORING, flags: if X_DROP { BA_XDROP } else { 0 }, matrix: matrix.as_ptr(), gaps,
                         size: SizeRange { min: *size.start(), max: *size.end() }, x_drop, cigar_eq: 0 };
    let mut out = vec![AlignResult::default(); n];
    let mut used = vec![0usize; n];
    let mut stats = BaStats::default();
    check(unsafe { ba_align_batch_exp(dev.h, &cfg, n, qa.as_ptr(), qo.as_ptr(), ra.as_ptr(), ro.as_ptr(), target_scores.as_ptr(),
                                      out.as_mut_ptr(), used.as_mut_ptr(), &mut stats) });
    (out, used.into_iter().map(|s| if s == 0 { None } else { Some(s) }).collect())
}

/// ASCII bases -> BAM nibble codes (two per byte) for `BA_INPUT_NUC4` batches; `Err(i)` = byte i is not an IUPAC base.
pub fn pack_nuc4(ascii: &[u8]) -> Result<Vec<u8>, usize> {
    let mut packed = vec![0u8; (ascii.len() + 1) / 2];
    match unsafe { ba_pack_nuc4(ascii.as_ptr(), ascii.len(), packed.as_mut_ptr(), 0) } { 0 => Ok(packed), i => Err(i - 1) }
}

#[cfg(test)]
mod tests {
    use super::*;
    #[test]
    fn doc_example() {   // the reference's doc-test, lib.rs:8-35
        let (min_block, max_block) = (32usize, 256usize);
        let r = PaddedBytes::from_bytes::<NucMatrix>(b"TTAAAAAAATTTTTTTTTTTT", max_block);
        let q = PaddedBytes::from_bytes::<NucMatrix>(b"TTTTTTTTAAAAAAATTTTTTTTT", max_block);
        let mut a = Block::<true, false>::new(q.len(), r.len(), max_block);
        a.align(&q, &r, nw1(), Gaps { open: -2, extend: -1 }, min_block..=max_block, 0);
        let res = a.res();
        assert_eq!(res, AlignResult { score: 7, query_idx: 24, reference_idx: 21 });
        let mut cigar = Cigar::new(res.query_idx, res.reference_idx);
        a.trace().cigar_eq(&q, &r, res.query_idx, res.reference_idx, &mut cigar);
        assert_eq!(cigar.to_string(), "2=6I16=3D");
    }
}
