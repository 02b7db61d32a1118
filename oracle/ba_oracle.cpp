// ba_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's AVX2 path (block-aligner v0.5.1 @ 4fcf630), written in C++
// with the *same* `_mm256_*` / `_mm_*` intrinsics the reference's `src/avx2.rs` uses, so that
// every lane of every vector evolves identically (garbage lanes and saturation included).
//
// Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
// `bench.py` may load this library; the product (`block_aligner_b200/csrc`) never links it.
//
// The original is Rust and cannot be compiled in this environment (no rustc/cargo), so parity of
// this file is pinned against every known-answer assertion in the reference's own tests
// (`src/scan_block.rs:1908-2230`, `src/avx2.rs:470-488`, `src/lib.rs:8-35`) -- see
// tests/test_oracle_golden.py.  The adaptive (grow/shrink) path has no golden vectors in the
// reference; there this file is cross-checked against an independently written scalar
// restatement (oracle/ba_scalar.cpp).
//
// Citations are `path:line` into /root/reference/src/.
//
// Build: see oracle/Makefile (g++ -O3 -mavx2 -shared -fPIC).

#include <immintrin.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "ba_oracle.h"

namespace {

// ---------------------------------------------------------------------------------------------
// avx2.rs restated (avx2.rs:6-16)
// ---------------------------------------------------------------------------------------------
typedef __m256i Simd;
typedef __m128i HalfSimd;
typedef int32_t TraceType;
constexpr size_t L = 16;
constexpr size_t L_BYTES = 32;
constexpr int16_t ZERO = 1 << 14;
constexpr int16_t MIN = 0;

constexpr size_t STEP = 8;               // scan_block.rs:787
constexpr size_t X_DROP_ITER = 2;        // scan_block.rs:788
constexpr bool SHRINK = true;            // scan_block.rs:789
constexpr size_t SHRINK_SUFFIX_LEN = STEP / 4;  // scan_block.rs:790

#define BA_INLINE static inline __attribute__((always_inline))
#define BA_MINLINE inline __attribute__((always_inline))

BA_INLINE Simd adds(Simd a, Simd b) { return _mm256_adds_epi16(a, b); }   // avx2.rs:26
BA_INLINE Simd subs(Simd a, Simd b) { return _mm256_subs_epi16(a, b); }   // avx2.rs:30
BA_INLINE Simd vmax(Simd a, Simd b) { return _mm256_max_epi16(a, b); }    // avx2.rs:34
BA_INLINE Simd cmpeq(Simd a, Simd b) { return _mm256_cmpeq_epi16(a, b); } // avx2.rs:38
BA_INLINE Simd blend8(Simd a, Simd b, Simd m) { return _mm256_blendv_epi8(a, b, m); } // avx2.rs:46
BA_INLINE Simd vload(const int16_t* p) { return _mm256_load_si256((const Simd*)p); }
BA_INLINE Simd vloadu(const int16_t* p) { return _mm256_loadu_si256((const Simd*)p); }
BA_INLINE void vstore(int16_t* p, Simd a) { _mm256_store_si256((Simd*)p, a); }
BA_INLINE Simd set1(int16_t v) { return _mm256_set1_epi16(v); }
BA_INLINE uint32_t movemask8(Simd a) { return (uint32_t)_mm256_movemask_epi8(a); } // avx2.rs:96

// avx2.rs:100-115: shift `a` up by one lane, pulling in the top lane of `b` at lane 0.
BA_INLINE Simd sl1(Simd a, Simd b) {
  return _mm256_alignr_epi8(a, _mm256_permute2x128_si256(a, b, 0x03), (int)(L_BYTES / 2 - 2));
}
// avx2.rs:139-141 (STEP = 8 lanes)
BA_INLINE Simd step8(Simd a, Simd b) { return _mm256_permute2x128_si256(a, b, 0x03); }
// avx2.rs:166-169
BA_INLINE Simd broadcasthi(Simd v) {
  v = _mm256_shufflehi_epi16(v, 0xFF);
  return _mm256_permute4x64_epi64(v, 0xFF);
}
// avx2.rs:173-182
BA_INLINE int16_t slow_extract(Simd v, size_t i) {
  alignas(32) int16_t a[L];
  vstore(a, v);
  return a[i];
}
// avx2.rs:186-192
BA_INLINE int16_t hmax(Simd v) {
  Simd v2 = _mm256_max_epi16(v, _mm256_srli_si256(v, 2));
  v2 = _mm256_max_epi16(v2, _mm256_srli_si256(v2, 4));
  v2 = _mm256_max_epi16(v2, _mm256_srli_si256(v2, 8));
  v2 = _mm256_max_epi16(v2, _mm256_permute2x128_si256(v2, v2, 0x03));
  return (int16_t)_mm256_extract_epi16(v2, 0);
}
// avx2.rs:221-242 with num = STEP = 8
BA_INLINE int16_t prefix_hmax8(Simd v) {
  v = _mm256_max_epi16(v, _mm256_srli_si256(v, 8));
  v = _mm256_max_epi16(v, _mm256_srli_si256(v, 4));
  v = _mm256_max_epi16(v, _mm256_srli_si256(v, 2));
  return (int16_t)_mm256_extract_epi16(v, 0);
}
// avx2.rs:246-267 with num = SHRINK_SUFFIX_LEN = 2
BA_INLINE int16_t suffix_hmax2(Simd v) {
  v = _mm256_max_epi16(v, _mm256_slli_si256(v, 2));
  return (int16_t)_mm256_extract_epi16(v, 15);
}
// avx2.rs:271-274
BA_INLINE size_t hargmax(Simd v, int16_t m) {
  Simd v2 = _mm256_cmpeq_epi16(v, _mm256_set1_epi16(m));
  return (size_t)__builtin_ctz(movemask8(v2)) / 2;
}
// avx2.rs:297-310
BA_INLINE void prefix_scan_consts(Simd gap, Simd* gap_all, Simd* consts) {
  Simd s1 = _mm256_slli_si256(gap, 2);
  s1 = _mm256_adds_epi16(s1, gap);
  Simd s2 = _mm256_slli_si256(s1, 4);
  s2 = _mm256_adds_epi16(s2, s1);
  Simd s4 = _mm256_slli_si256(s2, 8);
  s4 = _mm256_adds_epi16(s4, s2);
  Simd c1 = _mm256_srli_si256(_mm256_shufflehi_epi16(s4, 0xFF), 8);
  c1 = _mm256_permute4x64_epi64(c1, 0x05);
  c1 = _mm256_adds_epi16(c1, s4);
  *gap_all = c1;
  *consts = s4;
}
// avx2.rs:315-338
BA_INLINE Simd prefix_scan(Simd R_max, Simd gap_cost, Simd gap_cost_lane) {
  Simd s1 = _mm256_slli_si256(R_max, 2);
  s1 = _mm256_adds_epi16(s1, gap_cost);
  s1 = _mm256_max_epi16(R_max, s1);
  Simd s2 = _mm256_slli_si256(s1, 4);
  s2 = _mm256_adds_epi16(s2, _mm256_slli_epi16(gap_cost, 1));
  s2 = _mm256_max_epi16(s1, s2);
  Simd s4 = _mm256_slli_si256(s2, 8);
  s4 = _mm256_adds_epi16(s4, _mm256_slli_epi16(gap_cost, 2));
  s4 = _mm256_max_epi16(s2, s4);
  Simd c1 = _mm256_shufflehi_epi16(s4, 0xFF);
  c1 = _mm256_permute4x64_epi64(c1, 0x50);
  c1 = _mm256_adds_epi16(c1, gap_cost_lane);
  return _mm256_max_epi16(s4, c1);
}
// avx2.rs:343-350
BA_INLINE Simd lookup2(HalfSimd lut1, HalfSimd lut2, HalfSimd v) {
  HalfSimd a = _mm_shuffle_epi8(lut1, v);
  HalfSimd b = _mm_shuffle_epi8(lut2, v);
  HalfSimd mask = _mm_slli_epi16(v, 3);
  return _mm256_cvtepi8_epi16(_mm_blendv_epi8(a, b, mask));
}
// avx2.rs:354-356
BA_INLINE Simd lookup1(HalfSimd lut, HalfSimd v) { return _mm256_cvtepi8_epi16(_mm_shuffle_epi8(lut, v)); }
// avx2.rs:360-364
BA_INLINE Simd lookup_bytes(HalfSimd match, HalfSimd mismatch, HalfSimd a, HalfSimd b) {
  HalfSimd mask = _mm_cmpeq_epi8(a, b);
  return _mm256_cvtepi8_epi16(_mm_blendv_epi8(mismatch, match, mask));
}

BA_INLINE int16_t clamp16(int32_t x) {  // scan_block.rs:1704-1706
  return (int16_t)std::min(std::max(x, (int32_t)INT16_MIN), (int32_t)INT16_MAX);
}

// ---------------------------------------------------------------------------------------------
// scores.rs restated: three matrix kinds behind one get_scores() shape (scores.rs:121-127,
// 204-209, 263-267), plus the profile (scores.rs:454-715).
// ---------------------------------------------------------------------------------------------
struct NucM {
  const int8_t* t;  // 8*16, 16-byte aligned rows
  BA_MINLINE Simd scores(uint8_t c, HalfSimd v) const {
    return lookup1(_mm_loadu_si128((const HalfSimd*)(t + (size_t)(c & 7) * 16)), v);
  }
};
struct AAM {
  const int8_t* t;  // 27*32
  BA_MINLINE Simd scores(uint8_t c, HalfSimd v) const {
    const int8_t* p = t + (size_t)c * 32;
    return lookup2(_mm_loadu_si128((const HalfSimd*)p), _mm_loadu_si128((const HalfSimd*)(p + 16)), v);
  }
};
struct ByteM {
  int8_t match, mismatch;
  BA_MINLINE Simd scores(uint8_t c, HalfSimd v) const {
    return lookup_bytes(_mm_set1_epi8(match), _mm_set1_epi8(mismatch), _mm_set1_epi8((char)c), v);
  }
};

struct View {  // PaddedBytes as seen by the aligner (scan_block.rs:1790-1884): p[0] is the pad
  const uint8_t* p;
  size_t n;
  BA_MINLINE uint8_t get(size_t i) const { return p[i]; }
  BA_MINLINE const uint8_t* at(size_t i) const { return p + i; }
  BA_MINLINE size_t len() const { return n; }
};

}  // namespace

// scores.rs:454-468, 473-486
struct OraProfile {
  std::vector<int16_t> aa_pos;
  std::vector<int8_t> pos_aa;
  int8_t gap_extend;
  std::vector<int16_t> pos_gap_open_C, pos_gap_close_C, pos_gap_open_R;
  size_t max_len, curr_len, str_len;

  size_t len() const { return str_len; }
  BA_MINLINE Simd scores_pos(size_t i, HalfSimd v) const {  // scores.rs:596-602
    const int8_t* p = pos_aa.data() + i * 32;
    return lookup2(_mm_loadu_si128((const HalfSimd*)p), _mm_loadu_si128((const HalfSimd*)(p + 16)), v);
  }
  BA_MINLINE Simd scores_aa(size_t i, uint8_t c) const {  // scores.rs:609-612
    return vloadu(aa_pos.data() + (size_t)c * curr_len + i);
  }
};

namespace {

// ---------------------------------------------------------------------------------------------
// cigar.rs restated
// ---------------------------------------------------------------------------------------------
struct CigarBuf {
  std::vector<ora_oplen> s;
  size_t idx;
  void clear(size_t ql, size_t rl) {  // cigar.rs:58-61
    size_t n = ql + rl + 5;
    if (s.size() < n) s.resize(n);
    for (size_t k = 0; k < n; k++) { s[k].op = 0; s[k].len = 0; }
    idx = 1;
  }
  BA_MINLINE void add(uint8_t op) {  // cigar.rs:71-79
    size_t a = (op != s[idx - 1].op) ? 1 : 0;
    idx += a;
    s[idx - 1].op = op;
    s[idx - 1].len += 1;
  }
};

enum : uint8_t { OP_SENT = 0, OP_M = 1, OP_EQ = 2, OP_X = 3, OP_I = 4, OP_D = 5 };

// ---------------------------------------------------------------------------------------------
// Trace (scan_block.rs:1344-1692)
// ---------------------------------------------------------------------------------------------
struct LutEntry { uint8_t op, di, dj, table; };
enum : uint8_t { T_D = 0, T_C = 1, T_R = 2 };

static LutEntry g_lut[2][64];
static bool g_lut_init = false;
static void init_lut() {  // scan_block.rs:1508-1572
  if (g_lut_init) return;
  for (int right = 0; right < 2; right++)
    for (int tr = 0; tr < 4; tr++)
      for (int t2 = 0; t2 < 4; t2++) {
        // default entries (the reference fills the whole table with (D,0,1,D) first, including
        // the unused table index 3)
        for (int tb = 0; tb < 4; tb++) g_lut[right][(tr << 4) | (t2 << 2) | tb] = LutEntry{OP_D, 0, 1, T_D};
        for (int tb = 0; tb < 3; tb++) {
          LutEntry e;
          bool c_open_bit = (t2 & 1) != 0, r_open_bit = (t2 & 2) != 0;
          if (right) {
            if (tb == T_C) e = c_open_bit ? LutEntry{OP_D, 0, 1, T_D} : LutEntry{OP_D, 0, 1, T_C};
            else if (tb == T_R) e = r_open_bit ? LutEntry{OP_I, 1, 0, T_D} : LutEntry{OP_I, 1, 0, T_R};
            else if (tr == 0) e = LutEntry{OP_M, 1, 1, T_D};
            else if (tr & 1) e = c_open_bit ? LutEntry{OP_D, 0, 1, T_D} : LutEntry{OP_D, 0, 1, T_C};
            else e = r_open_bit ? LutEntry{OP_I, 1, 0, T_D} : LutEntry{OP_I, 1, 0, T_R};
          } else {
            // "everything is basically swapped" (scan_block.rs:1546-1558): bit0 of trace/trace2
            // now talks about the R table, bit1 about the C table.
            if (tb == T_R) e = c_open_bit ? LutEntry{OP_I, 1, 0, T_D} : LutEntry{OP_I, 1, 0, T_R};
            else if (tb == T_C) e = r_open_bit ? LutEntry{OP_D, 0, 1, T_D} : LutEntry{OP_D, 0, 1, T_C};
            else if (tr == 0) e = LutEntry{OP_M, 1, 1, T_D};
            else if (tr & 1) e = c_open_bit ? LutEntry{OP_I, 1, 0, T_D} : LutEntry{OP_I, 1, 0, T_R};
            else e = r_open_bit ? LutEntry{OP_D, 0, 1, T_D} : LutEntry{OP_D, 0, 1, T_C};
          }
          g_lut[right][(tr << 4) | (t2 << 2) | tb] = e;
        }
      }
  g_lut_init = true;
}

struct Trace {
  std::vector<TraceType> trace, trace2, zero_mask;
  std::vector<uint64_t> right;
  std::vector<uint32_t> block_start;
  std::vector<uint16_t> block_size;
  size_t trace_idx = 0, block_idx = 0, ckpt_trace_idx = 0, ckpt_block_idx = 0;
  size_t query_len = 0, reference_len = 0;
  bool local_start = false, free_query_start_gaps = false;

  void init(size_t ql, size_t rl, size_t max_size, bool local, bool fqs) {  // :1363-1392
    size_t len = ql + rl + 2;
    size_t n = (max_size / L) * (len + max_size * 2);
    trace.assign(n, 0);
    trace2.assign(n, 0);
    right.assign((len + 63) / 64, 0);
    block_start.assign(len * 2, 0);
    block_size.assign(len * 2, 0);
    if (local) zero_mask.assign(n, 0); else zero_mask.clear();
    trace_idx = block_idx = ckpt_trace_idx = ckpt_block_idx = 0;
    query_len = ql; reference_len = rl; local_start = local; free_query_start_gaps = fqs;
  }
  void clear(size_t ql, size_t rl) {  // :1395-1404
    std::fill(right.begin(), right.end(), 0);
    trace_idx = block_idx = ckpt_trace_idx = ckpt_block_idx = 0;
    query_len = ql; reference_len = rl;
  }
  BA_MINLINE void add_trace(TraceType t, TraceType t2) { trace[trace_idx] = t; trace2[trace_idx] = t2; trace_idx++; }
  BA_MINLINE void add_zero_mask(TraceType m) { zero_mask[trace_idx] = m; }
  BA_MINLINE void add_block(size_t i, size_t j, size_t width, size_t height, bool r) {  // :1428-1443
    block_start[block_idx * 2] = (uint32_t)i;
    block_start[block_idx * 2 + 1] = (uint32_t)j;
    block_size[block_idx * 2] = (uint16_t)height;
    block_size[block_idx * 2 + 1] = (uint16_t)width;
    size_t a = block_idx / 64, b = block_idx % 64;
    uint64_t v = right[a] & ~(1ull << b);
    right[a] = v | ((uint64_t)r << b);
    block_idx++;
  }
  BA_MINLINE void add_trace_idx(size_t a) { trace_idx += a; }
  BA_MINLINE void save_ckpt() { ckpt_trace_idx = trace_idx; ckpt_block_idx = block_idx; }
  BA_MINLINE void restore_ckpt() { trace_idx = ckpt_trace_idx; block_idx = ckpt_block_idx; }

  // scan_block.rs:1482-1672
  template <bool EQ>
  int cigar_core(size_t i, size_t j, const View* q, const View* r, CigarBuf& cigar) const {
    if (!(i <= query_len && j <= reference_len)) return -1;
    init_lut();
    cigar.clear(i, j);
    size_t bidx = block_idx, tidx = trace_idx;
    size_t block_i = 0, block_j = 0, block_w = 0, block_h = 0, rt = 0;
    uint8_t table = T_D;
    while (i > 0 || j > 0) {
      for (;;) {
        bidx -= 1;
        block_i = block_start[bidx * 2];
        block_j = block_start[bidx * 2 + 1];
        block_h = block_size[bidx * 2];
        block_w = block_size[bidx * 2 + 1];
        tidx -= block_w * block_h / L;
        if (i >= block_i && j >= block_j) {
          rt = (size_t)((right[bidx / 64] >> (bidx % 64)) & 1);
          break;
        }
      }
      const LutEntry* lut = g_lut[rt];
      if (rt) {
        while (i >= block_i && j >= block_j && (i > 0 || j > 0)) {
          if (free_query_start_gaps && i == 0) return 0;
          size_t ci = i - block_i, cj = j - block_j;
          size_t idx = tidx + ci / L + cj * (block_h / L);
          if (local_start && table == T_D) {
            if ((zero_mask[idx] >> ((ci % L) * 2)) & 1) return 0;
          }
          size_t t = ((uint32_t)trace[idx] >> ((ci % L) * 2)) & 3;
          size_t t2 = ((uint32_t)trace2[idx] >> ((ci % L) * 2)) & 3;
          const LutEntry& e = lut[(t << 4) | (t2 << 2) | table];
          uint8_t op = e.op;
          if (EQ && op == OP_M) op = (q->get(i) == r->get(j)) ? OP_EQ : OP_X;
          i -= e.di; j -= e.dj; table = e.table;
          cigar.add(op);
        }
      } else {
        while (i >= block_i && j >= block_j && (i > 0 || j > 0)) {
          size_t ci = i - block_i, cj = j - block_j;
          size_t idx = tidx + cj / L + ci * (block_w / L);
          if (local_start && table == T_D) {
            if ((zero_mask[idx] >> ((cj % L) * 2)) & 1) return 0;
          }
          size_t t = ((uint32_t)trace[idx] >> ((cj % L) * 2)) & 3;
          size_t t2 = ((uint32_t)trace2[idx] >> ((cj % L) * 2)) & 3;
          const LutEntry& e = lut[(t << 4) | (t2 << 2) | table];
          uint8_t op = e.op;
          if (EQ && op == OP_M) op = (q->get(i) == r->get(j)) ? OP_EQ : OP_X;
          i -= e.di; j -= e.dj; table = e.table;
          cigar.add(op);
        }
      }
    }
    return 0;
  }
};

// scan_block.rs:1714-1783
struct Aligned {
  int16_t* ptr = nullptr;
  size_t n = 0;
  void alloc(size_t k) { n = k; ptr = (int16_t*)aligned_alloc(L_BYTES, std::max<size_t>(k * 2, L_BYTES)); memset(ptr, 0, std::max<size_t>(k * 2, L_BYTES)); }
  ~Aligned() { free(ptr); }
  void clear(size_t k) { for (size_t i = 0; i < k; i += L) vstore(ptr + i, set1(MIN)); }
  BA_MINLINE void set_vec(const Aligned& o, size_t idx) { vstore(ptr + idx, vload(o.ptr + idx)); }
  BA_MINLINE void copy_vec(size_t new_idx, size_t idx) { vstore(ptr + new_idx, vload(ptr + idx)); }
};

enum Dir : int { DIR_RIGHT = 0, DIR_DOWN = 1, DIR_GROW = 2 };

constexpr int F_TRACE = 1, F_XDROP = 2, F_LOCAL = 4, F_FQS = 8, F_FQE = 16;

struct Gaps { int8_t open, extend; };

}  // namespace

// scan_block.rs:89-92, 1252-1340
struct OraBlock {
  int flags;
  ora_result res;
  Trace trace;
  Aligned D_col, C_col, D_row, R_row, D_col_ckpt, C_col_ckpt, D_row_ckpt, R_row_ckpt, temp_buf1, temp_buf2;
  size_t query_len, reference_len, max_size;
  // instrumentation (not in the reference): work done and optional step log
  uint64_t cells, steps;
  std::vector<ora_step>* step_log = nullptr;
  CigarBuf cigar;
};

namespace {

// ---------------------------------------------------------------------------------------------
// place_block (scan_block.rs:1083-1228), sequence-sequence
// ---------------------------------------------------------------------------------------------
template <int F, class M>
static Simd place_block(const M& matrix, Gaps gaps, const View& query, const View& reference, Trace& trace,
                        size_t start_i, size_t start_j, size_t width, size_t height,
                        int16_t* D_col, int16_t* C_col, int16_t* D_row, int16_t* R_row,
                        Simd D_corner, int16_t relative_zero, bool right, Simd* argmax_i, Simd* argmax_j) {
  constexpr bool TRACE = F & F_TRACE, X_DROP = F & F_XDROP, LOCAL_START = F & F_LOCAL,
                 FQS = F & F_FQS, FQE = F & F_FQE;
  const Simd gap_open = set1(gaps.open), gap_extend = set1(gaps.extend);
  Simd gap_extend_all, consts;
  prefix_scan_consts(gap_extend, &gap_extend_all, &consts);
  Simd D_max = set1(MIN), D_argmax_i = set1(0), D_argmax_j = set1(0);
  *argmax_i = D_argmax_i; *argmax_j = D_argmax_j;
  if (width == 0 || height == 0) return D_max;

  for (size_t j = 0; j < width; j++) {
    Simd R01 = set1(MIN), D11 = set1(MIN), R11 = set1(MIN), prev_trace_R = set1(0);
    const uint8_t c = reference.get(start_j + j);
    for (size_t i = 0; i < height; i += L) {
      const Simd D10 = vload(D_col + i), C10 = vload(C_col + i);
      const Simd D00 = sl1(D10, D_corner);
      D_corner = D10;
      const Simd scores = matrix.scores(c, _mm_loadu_si128((const HalfSimd*)query.at(start_i + i)));
      D11 = adds(D00, scores);
      if ((!LOCAL_START && start_i + i == 0 && start_j + j == 0) || (FQS && right && start_i + i == 0))
        D11 = _mm256_insert_epi16(D11, relative_zero, 0);
      if (LOCAL_START) D11 = vmax(D11, set1(relative_zero));
      const Simd C11_open = adds(D10, gap_open);
      const Simd C11 = vmax(adds(C10, gap_extend), C11_open);
      D11 = vmax(D11, C11);
      const Simd D11_open = adds(D11, subs(gap_open, gap_extend));
      R11 = prefix_scan(D11_open, gap_extend, consts);
      R11 = vmax(R11, adds(broadcasthi(R01), gap_extend_all));
      D11 = vmax(D11, R11);
      R01 = R11;
      if (TRACE) {
        const Simd tDC = cmpeq(D11, C11), tDR = cmpeq(D11, R11);
        const Simd mask = set1((int16_t)0xFF00);
        const uint32_t td = movemask8(blend8(tDC, tDR, mask));
        const Simd tmpR = cmpeq(R11, D11_open);
        const Simd tR = sl1(tmpR, prev_trace_R);
        const uint32_t td2 = movemask8(blend8(cmpeq(C11, C11_open), tR, mask));
        prev_trace_R = tmpR;
        if (LOCAL_START) trace.add_zero_mask((TraceType)movemask8(cmpeq(D11, set1(relative_zero))));
        trace.add_trace((TraceType)td, (TraceType)td2);
      }
      D_max = vmax(D_max, D11);
      if (X_DROP || (FQE && start_i + i + L > query.len())) {
        const Simd m = cmpeq(D_max, D11);
        D_argmax_i = blend8(D_argmax_i, set1((int16_t)i), m);
        D_argmax_j = blend8(D_argmax_j, set1((int16_t)j), m);
      }
      vstore(D_col + i, D11);
      vstore(C_col + i, C11);
    }
    D_corner = set1(MIN);
    D_row[j] = (int16_t)_mm256_extract_epi16(D11, 15);
    R_row[j] = (int16_t)_mm256_extract_epi16(R11, 15);
    if (!X_DROP && !FQE && start_i + height > query.len() && start_j + j >= reference.len()) {
      if (TRACE) trace.add_trace_idx((width - 1 - j) * (height / L));
      break;
    }
  }
  *argmax_i = D_argmax_i; *argmax_j = D_argmax_j;
  return D_max;
}

// ---------------------------------------------------------------------------------------------
// place_block_profile_{right,down} (scan_block.rs:612-783)
// RIGHT=true : vectors run along the query (rows), one profile position per column.
// RIGHT=false: vectors run along the profile positions, one query byte per "column".
// ---------------------------------------------------------------------------------------------
template <int F, bool RIGHT>
static Simd place_block_profile(const View& q, const OraProfile& r, Trace& trace,
                                size_t start_i, size_t start_j, size_t width, size_t height,
                                int16_t* D_col, int16_t* C_col, int16_t* D_row, int16_t* R_row,
                                Simd D_corner, int16_t relative_zero, Simd* argmax_i, Simd* argmax_j) {
  constexpr bool TRACE = F & F_TRACE, X_DROP = F & F_XDROP, LOCAL_START = F & F_LOCAL,
                 FQS = F & F_FQS, FQE = F & F_FQE;
  // In the macro, `$query`/`$reference` name the *vector-direction* / *column-direction* objects:
  // RIGHT: query = q (PaddedBytes), reference = profile.  DOWN: query = profile, reference = q.
  const size_t vec_len = RIGHT ? q.len() : r.len();   // `$query.len()`
  const size_t col_len = RIGHT ? r.len() : q.len();   // `$reference.len()`
  const Simd gap_extend = set1(r.gap_extend);
  Simd gap_extend_all, consts;
  prefix_scan_consts(gap_extend, &gap_extend_all, &consts);
  Simd D_max = set1(MIN), D_argmax_i = set1(0), D_argmax_j = set1(0);
  *argmax_i = D_argmax_i; *argmax_j = D_argmax_j;
  size_t idx = 0;
  Simd gap_open_C = set1(MIN), gap_close_C = set1(MIN), gap_open_R = set1(MIN), gap_close_R = set1(MIN);
  if (width == 0 || height == 0) return D_max;

  for (size_t j = 0; j < width; j++) {
    Simd R01 = set1(MIN), D11 = set1(MIN), R11 = set1(MIN), prev_trace_R = set1(0);
    if (RIGHT) {
      idx = start_j + j;
      gap_open_C = set1(r.pos_gap_open_C[idx]);
      gap_close_C = set1(r.pos_gap_close_C[idx]);
      gap_open_R = set1(r.pos_gap_open_R[idx]);
    }
    for (size_t i = 0; i < height; i += L) {
      const Simd D10 = vload(D_col + i), C10 = vload(C_col + i);
      const Simd D00 = sl1(D10, D_corner);
      D_corner = D10;
      if (!RIGHT) {
        idx = start_i + i;
        gap_open_C = vloadu(r.pos_gap_open_R.data() + idx);   // get_gap_open_down_R  (:673)
        gap_open_R = vloadu(r.pos_gap_open_C.data() + idx);   // get_gap_open_down_C  (:674)
        gap_close_R = vloadu(r.pos_gap_close_C.data() + idx); // get_gap_close_down_C (:675)
      }
      const Simd scores = RIGHT ? r.scores_pos(idx, _mm_loadu_si128((const HalfSimd*)q.at(start_i + i)))
                                : r.scores_aa(idx, q.get(start_j + j));
      D11 = adds(D00, scores);
      if ((!LOCAL_START && start_i + i == 0 && start_j + j == 0) || (FQS && RIGHT && start_i + i == 0))
        D11 = _mm256_insert_epi16(D11, relative_zero, 0);
      if (LOCAL_START) D11 = vmax(D11, set1(relative_zero));
      const Simd C11_open = adds(D10, adds(gap_open_C, gap_extend));
      const Simd C11 = vmax(adds(C10, gap_extend), C11_open);
      const Simd C11_end = RIGHT ? adds(C11, gap_close_C) : C11;
      D11 = vmax(D11, C11_end);
      const Simd D11_open = adds(D11, gap_open_R);
      R11 = prefix_scan(D11_open, gap_extend, consts);
      R11 = vmax(R11, adds(broadcasthi(R01), gap_extend_all));
      const Simd R11_end = RIGHT ? R11 : adds(R11, gap_close_R);
      D11 = vmax(D11, R11_end);
      R01 = R11;
      if (TRACE) {
        const Simd tDC = cmpeq(D11, C11_end), tDR = cmpeq(D11, R11_end);
        const Simd mask = set1((int16_t)0xFF00);
        const uint32_t td = movemask8(blend8(tDC, tDR, mask));
        const Simd tmpR = cmpeq(R11, D11_open);
        const Simd tR = sl1(tmpR, prev_trace_R);
        const uint32_t td2 = movemask8(blend8(cmpeq(C11, C11_open), tR, mask));
        prev_trace_R = tmpR;
        if (LOCAL_START) trace.add_zero_mask((TraceType)movemask8(cmpeq(D11, set1(relative_zero))));
        trace.add_trace((TraceType)td, (TraceType)td2);
      }
      D_max = vmax(D_max, D11);
      if (X_DROP || (FQE && start_i + i + L > vec_len)) {
        const Simd m = cmpeq(D_max, D11);
        D_argmax_i = blend8(D_argmax_i, set1((int16_t)i), m);
        D_argmax_j = blend8(D_argmax_j, set1((int16_t)j), m);
      }
      vstore(D_col + i, D11);
      vstore(C_col + i, C11);
    }
    D_corner = set1(MIN);
    D_row[j] = (int16_t)_mm256_extract_epi16(D11, 15);
    R_row[j] = (int16_t)_mm256_extract_epi16(R11, 15);
    if (!X_DROP && !FQE && start_i + height > vec_len && start_j + j >= col_len) {
      if (TRACE) trace.add_trace_idx((width - 1 - j) * (height / L));
      break;
    }
  }
  *argmax_i = D_argmax_i; *argmax_j = D_argmax_j;
  return D_max;
}

// helpers, scan_block.rs:1003-1061
BA_INLINE void just_offset(size_t bs, int16_t* b1, int16_t* b2, Simd off_add) {
  for (size_t i = 0; i < bs; i += L) {
    vstore(b1 + i, adds(vload(b1 + i), off_add));
    vstore(b2 + i, adds(vload(b2 + i), off_add));
  }
}
BA_INLINE int16_t prefix_max(const int16_t* buf) { return prefix_hmax8(vload(buf)); }
BA_INLINE int16_t suffix_max(const int16_t* buf, size_t len) { return suffix_hmax2(vload(buf + len - L)); }
BA_INLINE Simd shift_and_offset(size_t bs, int16_t* b1, int16_t* b2, const int16_t* t1, const int16_t* t2, Simd off_add) {
  Simd curr1 = adds(vload(b1), off_add);
  const Simd D_corner = set1((int16_t)_mm256_extract_epi16(curr1, (int)STEP - 1));
  Simd curr2 = adds(vload(b2), off_add);
  for (size_t i = 0; i + L < bs; i += L) {
    const Simd next1 = adds(vload(b1 + i + L), off_add);
    const Simd next2 = adds(vload(b2 + i + L), off_add);
    vstore(b1 + i, step8(next1, curr1));
    vstore(b2 + i, step8(next2, curr2));
    curr1 = next1;
    curr2 = next2;
  }
  vstore(b1 + bs - L, step8(vload(t1), curr1));
  vstore(b2 + bs - L, step8(vload(t2), curr2));
  return D_corner;
}

// A "scorer" bundles what differs between align_core and align_profile_core (scan_block.rs:994-995)
template <int F, class M>
struct SeqScorer {
  const M& matrix; Gaps gaps; View query, reference;
  size_t qlen() const { return query.len(); }
  size_t rlen() const { return reference.len(); }
  BA_MINLINE Simd right(Trace& t, size_t si, size_t sj, size_t w, size_t h, int16_t* a, int16_t* b, int16_t* c, int16_t* d,
                       Simd corner, int16_t rz, Simd* ai, Simd* aj) const {
    return place_block<F, M>(matrix, gaps, query, reference, t, si, sj, w, h, a, b, c, d, corner, rz, true, ai, aj);
  }
  BA_MINLINE Simd down(Trace& t, size_t si, size_t sj, size_t w, size_t h, int16_t* a, int16_t* b, int16_t* c, int16_t* d,
                      Simd corner, int16_t rz, Simd* ai, Simd* aj) const {
    return place_block<F, M>(matrix, gaps, reference, query, t, si, sj, w, h, a, b, c, d, corner, rz, false, ai, aj);
  }
};
template <int F>
struct ProfScorer {
  View query; const OraProfile& profile;
  size_t qlen() const { return query.len(); }
  size_t rlen() const { return profile.len(); }
  BA_MINLINE Simd right(Trace& t, size_t si, size_t sj, size_t w, size_t h, int16_t* a, int16_t* b, int16_t* c, int16_t* d,
                       Simd corner, int16_t rz, Simd* ai, Simd* aj) const {
    return place_block_profile<F, true>(query, profile, t, si, sj, w, h, a, b, c, d, corner, rz, ai, aj);
  }
  BA_MINLINE Simd down(Trace& t, size_t si, size_t sj, size_t w, size_t h, int16_t* a, int16_t* b, int16_t* c, int16_t* d,
                      Simd corner, int16_t rz, Simd* ai, Simd* aj) const {
    return place_block_profile<F, false>(query, profile, t, si, sj, w, h, a, b, c, d, corner, rz, ai, aj);
  }
};

// ---------------------------------------------------------------------------------------------
// align_core (scan_block.rs:94-595)
// ---------------------------------------------------------------------------------------------
template <int F, class S>
static void align_core(OraBlock& B, const S& sc, size_t min_size, size_t max_size, int32_t x_drop) {
  constexpr bool TRACE = F & F_TRACE, X_DROP = F & F_XDROP, FQE = F & F_FQE;
  const size_t qlen = sc.qlen(), rlen = sc.rlen();
  size_t si = 0, sj = 0;  // state.i, state.j
  int32_t best_max = 0;
  size_t best_argmax_i = 0, best_argmax_j = 0;
  int prev_dir = DIR_GROW, dir = DIR_GROW;
  size_t prev_size = 0, block_size = min_size;
  int32_t off = 0, prev_off, off_max = 0;
  size_t y_drop_iter = 0, x_drop_iter = 0;
  size_t i_ckpt = si, j_ckpt = sj;
  int32_t off_ckpt = 0;
  Simd D_corner = set1(MIN);
  Trace& tr = B.trace;
  int16_t *D_col = B.D_col.ptr, *C_col = B.C_col.ptr, *D_row = B.D_row.ptr, *R_row = B.R_row.ptr;
  int16_t *t1 = B.temp_buf1.ptr, *t2 = B.temp_buf2.ptr;

  for (;;) {
    prev_off = off;
    Simd grow_D_max = set1(MIN), grow_ai = set1(0), grow_aj = set1(0);
    Simd D_max, D_ai, D_aj;
    int16_t right_max, down_max;
    B.steps++;
    if (dir == DIR_RIGHT) {
      off = off_max;
      const Simd off_add = set1(clamp16(prev_off - off));
      if (TRACE) tr.add_block(si, sj + block_size - STEP, STEP, block_size, true);
      just_offset(block_size, D_col, C_col, off_add);
      B.cells += (uint64_t)STEP * block_size;
      D_max = sc.right(tr, si, sj + block_size - STEP, STEP, block_size, D_col, C_col, t1, t2,
                       prev_dir == DIR_DOWN ? adds(D_corner, off_add) : set1(MIN),
                       clamp16(-off + (int32_t)ZERO), &D_ai, &D_aj);
      right_max = prefix_max(D_col);
      D_corner = shift_and_offset(block_size, D_row, R_row, t1, t2, off_add);
      down_max = prefix_max(D_row);
    } else if (dir == DIR_DOWN) {
      off = off_max;
      const Simd off_add = set1(clamp16(prev_off - off));
      if (TRACE) tr.add_block(si + block_size - STEP, sj, block_size, STEP, false);
      just_offset(block_size, D_row, R_row, off_add);
      B.cells += (uint64_t)STEP * block_size;
      D_max = sc.down(tr, sj, si + block_size - STEP, STEP, block_size, D_row, R_row, t1, t2,
                      prev_dir == DIR_RIGHT ? adds(D_corner, off_add) : set1(MIN),
                      clamp16(-off + (int32_t)ZERO), &D_ai, &D_aj);
      down_max = prefix_max(D_row);
      D_corner = shift_and_offset(block_size, D_col, C_col, t1, t2, off_add);
      right_max = prefix_max(D_col);
    } else {
      D_corner = set1(MIN);
      const size_t grow_step = block_size - prev_size;
      if (TRACE) tr.add_block(si + prev_size, sj, prev_size, grow_step, false);
      B.cells += (uint64_t)grow_step * prev_size;
      Simd ai1, aj1;
      const Simd D_max1 = sc.down(tr, sj, si + prev_size, grow_step, prev_size, D_row, R_row,
                                  D_col + prev_size, C_col + prev_size, set1(MIN),
                                  clamp16(-off + (int32_t)ZERO), &ai1, &aj1);
      if (TRACE) tr.add_block(si, sj + prev_size, grow_step, block_size, true);
      B.cells += (uint64_t)grow_step * block_size;
      D_max = sc.right(tr, si, sj + prev_size, grow_step, block_size, D_col, C_col,
                       D_row + prev_size, R_row + prev_size, set1(MIN),
                       clamp16(-off + (int32_t)ZERO), &D_ai, &D_aj);
      right_max = prefix_max(D_col);
      down_max = prefix_max(D_row);
      grow_D_max = D_max1; grow_ai = ai1; grow_aj = aj1;
      for (size_t k = 0; k < block_size; k += L) {
        B.D_col_ckpt.set_vec(B.D_col, k); B.C_col_ckpt.set_vec(B.C_col, k);
        B.D_row_ckpt.set_vec(B.D_row, k); B.R_row_ckpt.set_vec(B.R_row, k);
      }
      if (TRACE) tr.save_ckpt();
    }

    prev_dir = dir;
    const int16_t D_max_max = FQE ? slow_extract(D_max, qlen % L) : hmax(D_max);
    const int16_t grow_max = hmax(grow_D_max);
    const int16_t max = std::max(D_max_max, grow_max);
    off_max = off + (int32_t)max - (int32_t)ZERO;
    y_drop_iter++;
    bool grow_no_max = dir == DIR_GROW;

    if (B.step_log) {
      ora_step s;
      s.dir = dir; s.i = (uint32_t)si; s.j = (uint32_t)sj; s.block_size = (uint32_t)block_size;
      s.off = off; s.max = max; s.right_max = right_max; s.down_max = down_max;
      B.step_log->push_back(s);
    }

    if (off_max > best_max) {
      if (FQE) {
        const size_t idx_j = (size_t)slow_extract(D_aj, qlen % L);
        best_argmax_i = qlen;
        if (dir == DIR_RIGHT) best_argmax_j = sj + (block_size - STEP) + idx_j;
        else if (dir == DIR_GROW) best_argmax_j = sj + prev_size + idx_j;
        else abort();  // unreachable!() in the reference
      }
      if (X_DROP) {
        const size_t lane_idx = hargmax(D_max, D_max_max);
        const size_t idx_i = (size_t)slow_extract(D_ai, lane_idx);
        const size_t idx_j = (size_t)slow_extract(D_aj, lane_idx);
        const size_t r = idx_i + lane_idx;
        const size_t c = (block_size - STEP) + idx_j;
        if (dir == DIR_RIGHT) { best_argmax_i = si + r; best_argmax_j = sj + c; }
        else if (dir == DIR_DOWN) { best_argmax_i = si + c; best_argmax_j = sj + r; }
        else {
          if (D_max_max >= grow_max) {
            best_argmax_i = si + idx_i + lane_idx;
            best_argmax_j = sj + prev_size + idx_j;
          } else {
            const size_t lane2 = hargmax(grow_D_max, grow_max);
            const size_t gi = (size_t)slow_extract(grow_ai, lane2);
            const size_t gj = (size_t)slow_extract(grow_aj, lane2);
            best_argmax_i = si + prev_size + gj;
            best_argmax_j = sj + gi + lane2;
          }
        }
      }
      if (block_size < max_size) {
        i_ckpt = si; j_ckpt = sj; off_ckpt = off;
        for (size_t k = 0; k < block_size; k += L) {
          B.D_col_ckpt.set_vec(B.D_col, k); B.C_col_ckpt.set_vec(B.C_col, k);
          B.D_row_ckpt.set_vec(B.D_row, k); B.R_row_ckpt.set_vec(B.R_row, k);
        }
        if (TRACE) tr.save_ckpt();
        grow_no_max = false;
      }
      best_max = off_max;
      y_drop_iter = 0;
    }

    if (X_DROP) {
      if (off_max < best_max - x_drop) {
        if (x_drop_iter < X_DROP_ITER - 1) x_drop_iter++;
        else break;
      } else {
        x_drop_iter = 0;
      }
    }

    if (si + block_size > qlen && sj + block_size > rlen) break;

    if (sj + block_size > rlen) { si += STEP; dir = DIR_DOWN; continue; }
    if (si + block_size > qlen) { sj += STEP; dir = DIR_RIGHT; continue; }

    const size_t next_size = block_size * 2;
    if (next_size <= max_size) {
      if (y_drop_iter > (block_size / STEP) - 1 || grow_no_max) {
        prev_size = block_size;
        block_size = next_size;
        dir = DIR_GROW;
        si = i_ckpt; sj = j_ckpt; off = off_ckpt;
        for (size_t k = 0; k < prev_size; k += L) {
          B.D_col.set_vec(B.D_col_ckpt, k); B.C_col.set_vec(B.C_col_ckpt, k);
          B.D_row.set_vec(B.D_row_ckpt, k); B.R_row.set_vec(B.R_row_ckpt, k);
        }
        if (TRACE) tr.restore_ckpt();
        y_drop_iter = 0;
        continue;
      }
    }

    if (SHRINK && block_size > min_size && y_drop_iter == 0) {
      const int16_t shrink_max = std::max(suffix_max(D_row, block_size), suffix_max(D_col, block_size));
      if (shrink_max >= max) {
        prev_dir = DIR_GROW;
        block_size /= 2;
        for (size_t k = 0; k < block_size; k += L) {
          B.D_col.copy_vec(k, k + block_size); B.C_col.copy_vec(k, k + block_size);
          B.D_row.copy_vec(k, k + block_size); B.R_row.copy_vec(k, k + block_size);
        }
        si += block_size; sj += block_size;
        i_ckpt = si; j_ckpt = sj; off_ckpt = off;
        for (size_t k = 0; k < block_size; k += L) {
          B.D_col_ckpt.set_vec(B.D_col, k); B.C_col_ckpt.set_vec(B.C_col, k);
          B.D_row_ckpt.set_vec(B.D_row, k); B.R_row_ckpt.set_vec(B.R_row, k);
        }
        right_max = prefix_max(D_col);
        down_max = prefix_max(D_row);
        if (TRACE) tr.save_ckpt();
        y_drop_iter = 0;
      }
    }

    if (down_max > right_max) { si += STEP; dir = DIR_DOWN; }
    else { sj += STEP; dir = DIR_RIGHT; }
  }

  if (X_DROP || FQE) {
    B.res.score = best_max; B.res.query_idx = best_argmax_i; B.res.reference_idx = best_argmax_j;
  } else {
    int32_t score;
    if (dir == DIR_RIGHT || dir == DIR_GROW) score = off + (int32_t)D_col[qlen - si] - (int32_t)ZERO;
    else score = off + (int32_t)D_row[rlen - sj] - (int32_t)ZERO;
    B.res.score = score; B.res.query_idx = qlen; B.res.reference_idx = rlen;
  }
}

static bool pow2(size_t x) { return x && !(x & (x - 1)); }

// Allocated::clear (scan_block.rs:1322-1339)
static int block_clear(OraBlock& B, size_t ql, size_t rl, size_t max_size) {
  if (!(ql + rl <= B.query_len + B.reference_len)) return ORA_ERR_TOO_LONG;
  if (!(max_size <= B.max_size)) return ORA_ERR_TOO_LONG;
  B.trace.clear(ql, rl);
  B.D_col.clear(max_size); B.C_col.clear(max_size); B.D_row.clear(max_size); B.R_row.clear(max_size);
  B.D_col_ckpt.clear(max_size); B.C_col_ckpt.clear(max_size); B.D_row_ckpt.clear(max_size); B.R_row_ckpt.clear(max_size);
  B.temp_buf1.clear(L); B.temp_buf2.clear(L);
  B.cells = 0; B.steps = 0;
  return 0;
}

template <int F, class M>
static void run_seq(OraBlock& B, const M& m, Gaps g, View q, View r, size_t mn, size_t mx, int32_t x) {
  SeqScorer<F, M> sc{m, g, q, r};
  align_core<F>(B, sc, mn, mx, x);
}
template <int F>
static void run_prof(OraBlock& B, View q, const OraProfile& p, size_t mn, size_t mx, int32_t x) {
  ProfScorer<F> sc{q, p};
  align_core<F>(B, sc, mn, mx, x);
}

// flags -> instantiation. Valid combos: TRACE x XDROP x {plain, LOCAL, FQS, FQE (no XDROP)}.
#define FOR_FLAGS(X) \
  X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(16) X(17) X(20) X(21) X(24) X(25)

template <class M>
static int dispatch_seq(OraBlock& B, const M& m, Gaps g, View q, View r, size_t mn, size_t mx, int32_t x) {
  switch (B.flags) {
#define X(FV) case FV: run_seq<FV, M>(B, m, g, q, r, mn, mx, x); return 0;
    FOR_FLAGS(X)
#undef X
  }
  return ORA_ERR_BAD_FLAGS;
}
static int dispatch_prof(OraBlock& B, View q, const OraProfile& p, size_t mn, size_t mx, int32_t x) {
  switch (B.flags) {
#define X(FV) case FV: run_prof<FV>(B, q, p, mn, mx, x); return 0;
    FOR_FLAGS(X)
#undef X
  }
  return ORA_ERR_BAD_FLAGS;
}

}  // namespace

// =================================================================================================
// C API
// =================================================================================================
extern "C" {

OraBlock* ora_block_new(size_t query_len, size_t reference_len, size_t max_size, int flags) {
  if (!pow2(max_size)) return nullptr;  // scan_block.rs:799
  OraBlock* B = new OraBlock();
  B->flags = flags;
  B->res = ora_result{0, 0, 0};
  if (flags & F_TRACE) B->trace.init(query_len, reference_len, max_size, flags & F_LOCAL, flags & F_FQS);
  else B->trace.init(0, 0, 0, false, false);
  for (Aligned* a : {&B->D_col, &B->C_col, &B->D_row, &B->R_row, &B->D_col_ckpt, &B->C_col_ckpt, &B->D_row_ckpt, &B->R_row_ckpt})
    a->alloc(max_size);
  B->temp_buf1.alloc(L); B->temp_buf2.alloc(L);
  B->query_len = query_len; B->reference_len = reference_len; B->max_size = max_size;
  B->cells = B->steps = 0;
  return B;
}
void ora_block_free(OraBlock* B) { delete B; }

// Block::align (scan_block.rs:847-878). q/r point at PaddedBytes storage (index 0 = pad byte).
int ora_align(OraBlock* B, const uint8_t* q, size_t qlen, const uint8_t* r, size_t rlen,
              int matrix_kind, const int8_t* matrix, int8_t gap_open, int8_t gap_extend,
              size_t min_size, size_t max_size, int32_t x_drop) {
  const int F = B->flags;
  if (!(gap_open < 0 && gap_extend < 0)) return ORA_ERR_GAPS;
  if (!(gap_open < gap_extend)) return ORA_ERR_GAPS;
  const size_t mn = min_size < L ? L : min_size, mx = max_size < L ? L : max_size;
  if (!(mn < 65535 && mx < 65535)) return ORA_ERR_SIZE;
  if (!(pow2(mn) && pow2(mx))) return ORA_ERR_SIZE;
  if ((F & F_XDROP) && x_drop < 0) return ORA_ERR_XDROP;
  if ((F & F_LOCAL) && (F & F_FQS)) return ORA_ERR_BAD_FLAGS;
  if ((F & F_XDROP) && (F & F_FQE)) return ORA_ERR_BAD_FLAGS;
  if ((F & F_FQE) && !(mn > qlen)) return ORA_ERR_SIZE;
  int e = block_clear(*B, qlen, rlen, mx);
  if (e) return e;
  View qv{q, qlen}, rv{r, rlen};
  Gaps g{gap_open, gap_extend};
  switch (matrix_kind) {
    case ORA_MATRIX_NUC: { NucM m{matrix}; return dispatch_seq(*B, m, g, qv, rv, mn, mx, x_drop); }
    case ORA_MATRIX_AA: { AAM m{matrix}; return dispatch_seq(*B, m, g, qv, rv, mn, mx, x_drop); }
    case ORA_MATRIX_BYTE: { ByteM m{matrix[0], matrix[1]}; return dispatch_seq(*B, m, g, qv, rv, mn, mx, x_drop); }
  }
  return ORA_ERR_BAD_FLAGS;
}

// Block::align_profile (scan_block.rs:942-968)
int ora_align_profile(OraBlock* B, const uint8_t* q, size_t qlen, const OraProfile* p,
                      size_t min_size, size_t max_size, int32_t x_drop) {
  const int F = B->flags;
  if (!(p->gap_extend < 0)) return ORA_ERR_GAPS;
  const size_t mn = min_size < L ? L : min_size, mx = max_size < L ? L : max_size;
  if (!(mn < 65535 && mx < 65535)) return ORA_ERR_SIZE;
  if (!(pow2(mn) && pow2(mx))) return ORA_ERR_SIZE;
  if ((F & F_XDROP) && x_drop < 0) return ORA_ERR_XDROP;
  if ((F & F_LOCAL) && (F & F_FQS)) return ORA_ERR_BAD_FLAGS;
  if ((F & F_XDROP) && (F & F_FQE)) return ORA_ERR_BAD_FLAGS;
  if ((F & F_FQE) && !(mn > qlen)) return ORA_ERR_SIZE;
  int e = block_clear(*B, qlen, p->len(), mx);
  if (e) return e;
  View qv{q, qlen};
  return dispatch_prof(*B, qv, *p, mn, mx, x_drop);
}

ora_result ora_res(const OraBlock* B) { return B->res; }
uint64_t ora_cells(const OraBlock* B) { return B->cells; }
uint64_t ora_steps(const OraBlock* B) { return B->steps; }

// Trace::cigar / cigar_eq (scan_block.rs:1469-1480). Writes ops in forward order.
// Returns the number of (op,len) runs, or -1 on a precondition failure.
long ora_cigar(OraBlock* B, size_t qi, size_t rj, int eq, const uint8_t* q, size_t qlen,
               const uint8_t* r, size_t rlen, ora_oplen* out, size_t cap) {
  if (!(B->flags & F_TRACE)) return -1;
  View qv{q, qlen}, rv{r, rlen};
  int e = eq ? B->trace.cigar_core<true>(qi, rj, &qv, &rv, B->cigar) : B->trace.cigar_core<false>(qi, rj, nullptr, nullptr, B->cigar);
  if (e) return -1;
  const size_t n = B->cigar.idx - 1;  // cigar.rs:87-89
  for (size_t k = 0; k < n && k < cap; k++) out[k] = B->cigar.s[B->cigar.idx - 1 - k];  // cigar.rs:92-94
  return (long)n;
}

// step log for --dump-steps style debugging
size_t ora_align_logged_steps(OraBlock* B, ora_step* out, size_t cap) {
  if (!B->step_log) return 0;
  size_t n = std::min(cap, B->step_log->size());
  memcpy(out, B->step_log->data(), n * sizeof(ora_step));
  return B->step_log->size();
}
void ora_enable_step_log(OraBlock* B, int on) {
  if (on && !B->step_log) B->step_log = new std::vector<ora_step>();
  if (B->step_log) B->step_log->clear();
  if (!on && B->step_log) { delete B->step_log; B->step_log = nullptr; }
}

// ---- AAProfile (scores.rs:470-715) ----
OraProfile* ora_profile_new(size_t str_len, size_t block_size, int8_t gap_extend) {
  OraProfile* p = new OraProfile();
  p->max_len = str_len + block_size + 1;
  p->aa_pos.assign(32 * p->max_len, (int16_t)INT8_MIN);
  p->pos_aa.assign(p->max_len * 32, INT8_MIN);
  p->gap_extend = gap_extend;
  p->pos_gap_open_C.assign(p->max_len, (int16_t)INT8_MIN);
  p->pos_gap_close_C.assign(p->max_len, (int16_t)INT8_MIN);
  p->pos_gap_open_R.assign(p->max_len, (int16_t)INT8_MIN);
  // the "down" path reads 16-lane vectors starting at idx (scores.rs:646-666, 609-612): keep
  // slack so those unaligned loads stay in bounds exactly as the Rust Vec capacity would allow
  p->curr_len = p->max_len;
  p->str_len = str_len;
  return p;
}
void ora_profile_free(OraProfile* p) { delete p; }
size_t ora_profile_len(const OraProfile* p) { return p->str_len; }
int ora_profile_clear(OraProfile* p, size_t str_len, size_t block_size) {
  size_t cl = str_len + block_size + 1;
  if (cl > p->max_len) return ORA_ERR_TOO_LONG;
  std::fill(p->aa_pos.begin(), p->aa_pos.begin() + 32 * cl, (int16_t)INT8_MIN);
  std::fill(p->pos_aa.begin(), p->pos_aa.begin() + cl * 32, INT8_MIN);
  std::fill(p->pos_gap_open_C.begin(), p->pos_gap_open_C.begin() + cl, (int16_t)INT8_MIN);
  std::fill(p->pos_gap_close_C.begin(), p->pos_gap_close_C.begin() + cl, (int16_t)INT8_MIN);
  std::fill(p->pos_gap_open_R.begin(), p->pos_gap_open_R.begin() + cl, (int16_t)INT8_MIN);
  p->str_len = str_len; p->curr_len = cl;
  return 0;
}
static inline uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
int ora_profile_set(OraProfile* p, size_t i, uint8_t b, int8_t score) {
  b = up(b);
  if (!(b >= 'A' && b <= 'Z' + 1)) return ORA_ERR_CHAR;
  p->pos_aa[i * 32 + (b - 'A')] = score;
  p->aa_pos[(size_t)(b - 'A') * p->curr_len + i] = score;
  return 0;
}
int ora_profile_set_all(OraProfile* p, const uint8_t* order, size_t order_len, const int8_t* scores,
                        size_t scores_len, size_t left_shift, size_t right_shift, int rev) {
  uint8_t o[32];
  for (int k = 0; k < 32; k++) o[k] = 26;
  if (order_len > 32) return ORA_ERR_SIZE;
  for (size_t k = 0; k < order_len; k++) {
    uint8_t b = up(order[k]);
    if (!(b >= 'A' && b <= 'Z' + 1)) return ORA_ERR_CHAR;
    o[k] = b - 'A';
  }
  if (scores_len / order_len != p->str_len) return ORA_ERR_SIZE;
  size_t score_idx = 0;
  for (size_t n = 0; n < p->str_len; n++) {
    size_t i = rev ? p->str_len - n : 1 + n;
    for (size_t j = 0; j < order_len; j++) {
      int8_t s = (int8_t)((int8_t)(scores[score_idx] << left_shift) >> right_shift);
      size_t b = o[j];
      p->pos_aa[i * 32 + b] = s;
      p->aa_pos[b * p->curr_len + i] = s;
      score_idx++;
    }
  }
  return 0;
}
int ora_profile_set_gap_open_C(OraProfile* p, size_t i, int8_t g) { if (!(g < 0)) return ORA_ERR_GAPS; p->pos_gap_open_C[i] = g; return 0; }
int ora_profile_set_gap_close_C(OraProfile* p, size_t i, int8_t g) { p->pos_gap_close_C[i] = g; return 0; }
int ora_profile_set_gap_open_R(OraProfile* p, size_t i, int8_t g) { if (!(g < 0)) return ORA_ERR_GAPS; p->pos_gap_open_R[i] = g; return 0; }
int ora_profile_set_all_gap_open_C(OraProfile* p, int8_t g) { if (!(g < 0)) return ORA_ERR_GAPS; std::fill(p->pos_gap_open_C.begin(), p->pos_gap_open_C.begin() + p->str_len + 1, (int16_t)g); return 0; }
int ora_profile_set_all_gap_close_C(OraProfile* p, int8_t g) { std::fill(p->pos_gap_close_C.begin(), p->pos_gap_close_C.begin() + p->str_len + 1, (int16_t)g); return 0; }
int ora_profile_set_all_gap_open_R(OraProfile* p, int8_t g) { if (!(g < 0)) return ORA_ERR_GAPS; std::fill(p->pos_gap_open_R.begin(), p->pos_gap_open_R.begin() + p->str_len + 1, (int16_t)g); return 0; }
int8_t ora_profile_get(const OraProfile* p, size_t i, uint8_t b) { return p->pos_aa[i * 32 + (up(b) - 'A')]; }
int8_t ora_profile_get_gap_extend(const OraProfile* p) { return p->gap_extend; }
// AAProfile::from_bytes (scores.rs:489-505)
OraProfile* ora_profile_from_bytes(const uint8_t* b, size_t len, size_t block_size, int8_t match, int8_t mismatch,
                                   int8_t gap_open_C, int8_t gap_close_C, int8_t gap_open_R, int8_t gap_extend) {
  OraProfile* p = ora_profile_new(len, block_size, gap_extend);
  for (size_t i = 0; i < len; i++)
    for (uint8_t c = 'A'; c <= 'Z'; c++) ora_profile_set(p, i + 1, c, c == b[i] ? match : mismatch);
  for (size_t i = 0; i < len + 1; i++) {
    ora_profile_set_gap_open_C(p, i, gap_open_C);
    ora_profile_set_gap_close_C(p, i, gap_close_C);
    ora_profile_set_gap_open_R(p, i, gap_open_R);
  }
  return p;
}
// raw views for tests / the GPU marshaller cross-check
const int8_t* ora_profile_pos_aa(const OraProfile* p) { return p->pos_aa.data(); }
const int16_t* ora_profile_gap_open_C(const OraProfile* p) { return p->pos_gap_open_C.data(); }
const int16_t* ora_profile_gap_close_C(const OraProfile* p) { return p->pos_gap_close_C.data(); }
const int16_t* ora_profile_gap_open_R(const OraProfile* p) { return p->pos_gap_open_R.data(); }
size_t ora_profile_curr_len(const OraProfile* p) { return p->curr_len; }

// ---- matrices (scores.rs:42-217) and PaddedBytes (scan_block.rs:1790-1884) ----
void ora_nuc_matrix_simple(int8_t match, int8_t mismatch, int8_t* out128) {  // scores.rs:150-164
  for (int k = 0; k < 128; k++) out128[k] = INT8_MIN;
  const uint8_t alpha[5] = {'A', 'T', 'C', 'G', 'N'};
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) out128[(alpha[i] & 7) * 16 + (alpha[j] & 15)] = (i == j) ? match : mismatch;
}
void ora_aa_matrix_simple(int8_t match, int8_t mismatch, int8_t* out864) {  // scores.rs:48-61
  for (int k = 0; k < 27 * 32; k++) out864[k] = INT8_MIN;
  for (int i = 0; i < 26; i++)
    for (int j = 0; j < 26; j++) out864[i * 32 + j] = (i == j) ? match : mismatch;
}
// writes 1 + len + block_size bytes; kind selects convert_char / NULL. Returns 0 or an error.
int ora_pad(int matrix_kind, const uint8_t* s, size_t len, size_t block_size, int rev, uint8_t* out) {
  uint8_t null_c = matrix_kind == ORA_MATRIX_NUC ? (uint8_t)'Z' : (matrix_kind == ORA_MATRIX_AA ? (uint8_t)26 : (uint8_t)0);
  out[0] = null_c;
  for (size_t k = 0; k < len; k++) {
    uint8_t c = s[rev ? len - 1 - k : k];
    if (matrix_kind == ORA_MATRIX_NUC) {
      c = up(c);
      if (!(c >= 'A' && c <= 'Z')) return ORA_ERR_CHAR;
    } else if (matrix_kind == ORA_MATRIX_AA) {
      c = up(c);
      if (!(c >= 'A' && c <= 'A' + 26)) return ORA_ERR_CHAR;
      c = c - 'A';
    }
    out[1 + k] = c;
  }
  for (size_t k = 0; k < block_size; k++) out[1 + len + k] = null_c;
  return 0;
}

// the prefix scan on its own, for the avx2.rs:470-488 golden vectors and for cross-checking the
// closed form used by the scalar restatement / CUDA kernel
void ora_prefix_scan(const int16_t* in16, int16_t gap, int16_t* out16) {
  alignas(32) int16_t a[16], o[16];
  memcpy(a, in16, 32);
  Simd g = set1(gap), all, consts;
  prefix_scan_consts(g, &all, &consts);
  vstore(o, prefix_scan(vload(a), g, consts));
  memcpy(out16, o, 32);
}
void ora_prefix_scan_consts(int16_t gap, int16_t* gap_all16, int16_t* consts16) {
  alignas(32) int16_t a[16], b[16];
  Simd all, consts;
  prefix_scan_consts(set1(gap), &all, &consts);
  vstore(a, all); vstore(b, consts);
  memcpy(gap_all16, a, 32); memcpy(consts16, b, 32);
}

static double g_last_batch_seconds = 0;
// ---- batch driver: the CPU baseline: one reusable Block per thread ----
// Sequences arrive already padded (see ora_pad) inside one arena; pair k uses q at q_off[k]
// (pad byte included) with length q_len[k], same for r. For profile batches `profiles[k]`
// replaces r. Results/cigars are written per pair. Returns 0 or the first error.
int ora_batch_align(const ora_batch* b, int n_threads, ora_result* out, uint64_t* out_cells,
                    ora_oplen* cigar_arena, const uint64_t* cigar_off, uint32_t* cigar_len) {
  if (n_threads < 1) n_threads = 1;
  size_t max_q = 0, max_r = 0;
  for (size_t k = 0; k < b->n; k++) {
    max_q = std::max<size_t>(max_q, b->q_len[k]);
    size_t rl = b->profiles ? b->profiles[k]->str_len : b->r_len[k];
    max_r = std::max<size_t>(max_r, rl);
  }
  const size_t mx = b->max_size < L ? L : b->max_size;
  std::atomic<int> err{0};
  auto worker = [&](int t) {
    OraBlock* B = ora_block_new(max_q, max_r, mx, b->flags);
    if (!B) { err = ORA_ERR_SIZE; return; }
    const size_t lo = b->n * (size_t)t / n_threads, hi = b->n * (size_t)(t + 1) / n_threads;
    for (size_t k = lo; k < hi; k++) {
      const uint8_t* q = b->arena + b->q_off[k];
      int e;
      const uint8_t* r = nullptr; size_t rl = 0;
      if (b->profiles) {
        e = ora_align_profile(B, q, b->q_len[k], b->profiles[k], b->min_size, b->max_size, b->x_drop);
        rl = b->profiles[k]->str_len;
      } else {
        r = b->arena + b->r_off[k]; rl = b->r_len[k];
        e = ora_align(B, q, b->q_len[k], r, rl, b->matrix_kind, b->matrix, b->gap_open, b->gap_extend,
                      b->min_size, b->max_size, b->x_drop);
      }
      if (e) { err = e; break; }
      out[k] = B->res;
      if (out_cells) out_cells[k] = B->cells;
      if (cigar_arena && (b->flags & F_TRACE)) {
        long n = ora_cigar(B, B->res.query_idx, B->res.reference_idx, b->cigar_eq && r, q, b->q_len[k], r, rl,
                           cigar_arena + cigar_off[k], (size_t)(cigar_off[k + 1] - cigar_off[k]));
        cigar_len[k] = n < 0 ? 0 : (uint32_t)n;
      }
    }
    ora_block_free(B);
  };
  // wall clock around the align (+cigar) calls only; Block allocation is inside too but is amortised
  // over the thread's share exactly like the reference's documented reuse pattern (scan_block.rs:795-797)
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(worker, t);
  worker(0);
  for (auto& t : th) t.join();
  g_last_batch_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return err.load();
}

// Pads a whole arena of raw sequences (PaddedBytes::from_bytes per sequence). `out_off` (n entries) receives the
// start of each padded sequence inside `out`, which must hold sum(1 + len + block_size) bytes.
int ora_pad_batch(int matrix_kind, const uint8_t* raw, const uint64_t* off, size_t n, size_t block_size, uint8_t* out,
                  uint64_t* out_off, uint32_t* out_len) {
  uint64_t pos = 0;
  for (size_t k = 0; k < n; k++) {
    const size_t len = (size_t)(off[k + 1] - off[k]);
    int e = ora_pad(matrix_kind, raw + off[k], len, block_size, 0, out + pos);
    if (e) return e;
    out_off[k] = pos; out_len[k] = (uint32_t)len;
    pos += 1 + len + block_size;
  }
  return 0;
}

double ora_last_batch_seconds(void) { return g_last_batch_seconds; }

int ora_hw_threads(void) { unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }

}  // extern "C"
