// block_aligner_b200.hpp -- the reference's Rust surface as header-only C++ over the C ABI ("Rust surface").
//
// Same names, argument order and meaning as `block_aligner::{scan_block, scores, cigar}`:
//
//     using namespace block_aligner;
//     auto r = PaddedBytes::from_bytes<NucMatrix>("TTAAAAAAATTTTTTTTTTTT", 16);          // README.md:32-59
//     auto q = PaddedBytes::from_bytes<NucMatrix>("TTTTTTTTAAAAAAATTTTTTTTT", 16);
//     Block<true, false> a(q.len(), r.len(), 16);                                        // Block::<TRACE, X_DROP>::new
//     a.align(q, r, NW1(), Gaps{-2, -1}, {16, 16}, 0);
//     AlignResult res = a.res();                                                         // {7, 24, 21}
//     Cigar cigar(res.query_idx, res.reference_idx);
//     a.trace().cigar_eq(q, r, res.query_idx, res.reference_idx, cigar);                 // "2=6I16=3D"
//
// Differences a Rust user will notice: `new` is the constructor; ranges are `SizeRange{min, max}`; statics are
// functions (`NW1()`, `BLOSUM62()`); precondition violations throw block_aligner::Error with the reference's
// panic message instead of aborting. Every alignment runs on the GPU through the batch C ABI
// (include/block_aligner_b200.h, Part 2) with a batch of one; for throughput use `align_batch` below or the
// C batch calls directly. Reference citations: src/scan_block.rs, src/scores.rs, src/cigar.rs, src/lib.rs.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

// The C ABI is kept inside block_aligner::c_abi so that its global names (AAMatrix, BLOSUM62, the Operation
// enumerators M / X / I / D ...) do not collide with the C++ types of the same name below.
#ifdef BLOCK_ALIGNER_B200_H
#error "include block_aligner_b200.hpp instead of (not after) block_aligner_b200.h"
#endif
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
namespace block_aligner { namespace c_abi {
#include "block_aligner_b200.h"
}}

namespace block_aligner {

using c_abi::AlignResult;   // scan_block.rs:1887-1893
using c_abi::Gaps;          // scores.rs:335-338
using c_abi::SizeRange;     // ffi.rs:20-23 (min..=max)

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

inline bool operator==(const AlignResult& a, const AlignResult& b) {
  return a.score == b.score && a.query_idx == b.query_idx && a.reference_idx == b.reference_idx;
}

// cigar.rs:17-39
enum class Operation : uint8_t { Sentinel = 0, M = 1, Eq = 2, X = 3, I = 4, D = 5 };
struct OpLen { Operation op; size_t len; };
inline bool operator==(const OpLen& a, const OpLen& b) { return a.op == b.op && a.len == b.len; }

namespace detail {
inline void check(int rc) {
  if (rc != c_abi::BA_OK) throw Error(std::string(c_abi::ba_error_string(rc)) + ": " + c_abi::ba_last_error_message());
}
inline void require(bool ok, const char* msg) { if (!ok) throw Error(msg); }
// one aligner (device $BA_DEVICE or 0) shared by every Block of the process, created on first use
inline c_abi::BaAligner* aligner() {
  static c_abi::BaAligner* a = [] {
    c_abi::BaAligner* p = nullptr;
    int dev = 0;
    if (const char* e = std::getenv("BA_DEVICE")) dev = std::atoi(e);
    check(c_abi::ba_create(dev, &p));
    return p;
  }();
  return a;
}
inline uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
}  // namespace detail

// ---- scoring matrices (scores.rs:40-273) -------------------------------------------------------------------
// Value types with the reference's `repr(C, align(32))` layouts, so the library's statics can be viewed in place.
struct alignas(32) AAMatrix {
  int8_t scores[27 * 32];
  static constexpr int kScoring = c_abi::BA_SCORING_AA;
  static constexpr uint8_t NULL_BYTE = 'A' + 26;                                  // scores.rs:83
  static AAMatrix new_() { AAMatrix m; std::memset(m.scores, 0x80, sizeof(m.scores)); return m; }   // scores.rs:48-50
  static AAMatrix new_simple(int8_t match_score, int8_t mismatch_score) {         // scores.rs:53-61
    AAMatrix m = new_();
    for (int a = 0; a < 26; a++) for (int b = 0; b < 26; b++) m.scores[a * 32 + b] = a == b ? match_score : mismatch_score;
    return m;
  }
  static uint8_t convert_char(uint8_t c) {                                        // scores.rs:130-134
    c = detail::upper(c);
    detail::require(c >= 'A' && c <= NULL_BYTE, "AAMatrix: byte outside A..=[");
    return (uint8_t)(c - 'A');
  }
  void set(uint8_t a, uint8_t b, int8_t score) {                                  // scores.rs:89-98
    const uint8_t x = convert_char(a), y = convert_char(b);
    scores[x * 32 + y] = score; scores[y * 32 + x] = score;
  }
  int8_t get(uint8_t a, uint8_t b) const { return scores[convert_char(a) * 32 + convert_char(b)]; }   // scores.rs:100-106
  const void* data() const { return scores; }
};
struct alignas(32) NucMatrix {
  int8_t scores[8 * 16];
  static constexpr int kScoring = c_abi::BA_SCORING_NUC;
  static constexpr uint8_t NULL_BYTE = 'Z';                                       // scores.rs:168
  static NucMatrix new_() { NucMatrix m; std::memset(m.scores, 0x80, sizeof(m.scores)); return m; }
  static NucMatrix new_simple(int8_t match_score, int8_t mismatch_score) {        // scores.rs:150-164
    NucMatrix m = new_();
    const char* nuc = "ACGNT";
    for (const char* a = nuc; *a; a++) for (const char* b = nuc; *b; b++) m.set((uint8_t)*a, (uint8_t)*b, *a == *b ? match_score : mismatch_score);
    return m;
  }
  static uint8_t convert_char(uint8_t c) {                                        // scores.rs:212-216
    c = detail::upper(c);
    detail::require(c >= 'A' && c <= 'Z', "NucMatrix: byte outside A..=Z");
    return c;
  }
  void set(uint8_t a, uint8_t b, int8_t score) {                                  // scores.rs:174-183
    const uint8_t x = convert_char(a), y = convert_char(b);
    scores[(x & 7) * 16 + (y & 15)] = score; scores[(y & 7) * 16 + (x & 15)] = score;
  }
  int8_t get(uint8_t a, uint8_t b) const { return scores[(convert_char(a) & 7) * 16 + (convert_char(b) & 15)]; }
  const void* data() const { return scores; }
};
struct ByteMatrix {                                                               // scores.rs:222-273
  int8_t match_score, mismatch_score;
  static constexpr int kScoring = c_abi::BA_SCORING_BYTE;
  static constexpr uint8_t NULL_BYTE = 0;
  static ByteMatrix new_simple(int8_t match_score, int8_t mismatch_score) { return ByteMatrix{match_score, mismatch_score}; }
  static uint8_t convert_char(uint8_t c) { return c; }
  int8_t get(uint8_t a, uint8_t b) const { return a == b ? match_score : mismatch_score; }
  const void* data() const { return &match_score; }
};

// the crate's statics (scores.rs:275-311; c/block_aligner.h:140-162), viewed in place
#define BA_HPP_STATIC_AA(NAME) inline const AAMatrix& NAME() { return *reinterpret_cast<const AAMatrix*>(&c_abi::NAME); }
BA_HPP_STATIC_AA(BLOSUM45) BA_HPP_STATIC_AA(BLOSUM50) BA_HPP_STATIC_AA(BLOSUM62) BA_HPP_STATIC_AA(BLOSUM80) BA_HPP_STATIC_AA(BLOSUM90)
BA_HPP_STATIC_AA(PAM100) BA_HPP_STATIC_AA(PAM120) BA_HPP_STATIC_AA(PAM160) BA_HPP_STATIC_AA(PAM200) BA_HPP_STATIC_AA(PAM250)
#undef BA_HPP_STATIC_AA
inline const NucMatrix& NW1() { return *reinterpret_cast<const NucMatrix*>(&c_abi::NW1); }
inline const ByteMatrix& BYTES1() { return *reinterpret_cast<const ByteMatrix*>(&c_abi::BYTES1); }

// lib.rs:109-111
inline size_t percent_len(size_t len, float p) { return c_abi::ba_percent_len(len, p); }

// ---- PaddedBytes (scan_block.rs:1790-1884) -----------------------------------------------------------------
class PaddedBytes {
 public:
  template <class M> static PaddedBytes new_(size_t len, size_t block_size) {       // :1798-1804
    PaddedBytes p;
    p.s_.assign(1 + len + block_size, M::NULL_BYTE == 'A' + 26 ? (uint8_t)26 : M::NULL_BYTE);
    p.len_ = len;
    return p;
  }
  template <class M> static PaddedBytes from_bytes(std::string_view b, size_t block_size) {   // :1829-1836
    PaddedBytes p = new_<M>(b.size(), block_size);
    p.set_bytes<M>(b, block_size);
    return p;
  }
  template <class M> static PaddedBytes from_str(std::string_view s, size_t block_size) { return from_bytes<M>(s, block_size); }
  template <class M> static PaddedBytes from_string(const std::string& s, size_t block_size) { return from_bytes<M>(s, block_size); }
  template <class M> void set_bytes(std::string_view b, size_t block_size) { fill<M>(b, block_size, false); }       // :1807-1813
  template <class M> void set_bytes_rev(std::string_view b, size_t block_size) { fill<M>(b, block_size, true); }    // :1815-1822
  uint8_t get(size_t i) const { return s_.at(i); }                                  // :1863-1866 (index 0 is the pad)
  size_t len() const { return len_; }
  const std::vector<uint8_t>& raw() const { return raw_; }                          // unconverted bytes, as aligned
 private:
  template <class M> void fill(std::string_view b, size_t block_size, bool rev) {
    detail::require(1 + b.size() + block_size <= s_.size(), "PaddedBytes: sequence does not fit the allocated length");
    const uint8_t nul = M::convert_char(M::NULL_BYTE);
    raw_.resize(b.size());
    s_[0] = nul;
    for (size_t k = 0; k < b.size(); k++) {
      const uint8_t c = (uint8_t)b[rev ? b.size() - 1 - k : k];
      raw_[k] = c;
      s_[1 + k] = M::convert_char(c);
    }
    for (size_t k = 0; k < block_size; k++) s_[1 + b.size() + k] = nul;
    len_ = b.size();
  }
  std::vector<uint8_t> s_, raw_;
  size_t len_ = 0;
};

// ---- AAProfile (scores.rs:454-715) over the C handle ----------------------------------------------------------
class AAProfile {
 public:
  AAProfile(size_t str_len, size_t block_size, int8_t gap_extend) : p_(c_abi::block_new_aaprofile(str_len, block_size, gap_extend)) {}
  static AAProfile from_bytes(std::string_view b, size_t block_size, int8_t match_score, int8_t mismatch_score,
                              int8_t gap_open_C, int8_t gap_close_C, int8_t gap_open_R, int8_t gap_extend) {   // :488-504
    AAProfile res(b.size(), block_size, gap_extend);
    for (size_t i = 0; i < b.size(); i++)
      for (uint8_t c = 'A'; c <= 'Z'; c++) res.set(i + 1, c, c == (uint8_t)b[i] ? match_score : mismatch_score);
    for (size_t i = 0; i < b.size() + 1; i++) { res.set_gap_open_C(i, gap_open_C); res.set_gap_close_C(i, gap_close_C); res.set_gap_open_R(i, gap_open_R); }
    return res;
  }
  AAProfile(AAProfile&& o) noexcept : p_(o.p_) { o.p_ = nullptr; }
  AAProfile& operator=(AAProfile&& o) noexcept { std::swap(p_, o.p_); return *this; }
  AAProfile(const AAProfile&) = delete;
  AAProfile& operator=(const AAProfile&) = delete;
  ~AAProfile() { if (p_) c_abi::block_free_aaprofile(p_); }
  size_t len() const { return c_abi::block_len_aaprofile(p_); }
  void clear(size_t str_len, size_t block_size) { c_abi::block_clear_aaprofile(p_, str_len, block_size); }
  void set(size_t i, uint8_t b, int8_t score) { c_abi::block_set_aaprofile(p_, i, b, score); }
  void set_all(std::string_view order, const std::vector<int8_t>& scores, size_t left_shift = 0, size_t right_shift = 0) {
    c_abi::block_set_all_aaprofile(p_, (const uint8_t*)order.data(), order.size(), scores.data(), scores.size(), left_shift, right_shift);
  }
  void set_all_rev(std::string_view order, const std::vector<int8_t>& scores, size_t left_shift = 0, size_t right_shift = 0) {
    c_abi::block_set_all_rev_aaprofile(p_, (const uint8_t*)order.data(), order.size(), scores.data(), scores.size(), left_shift, right_shift);
  }
  void set_gap_open_C(size_t i, int8_t gap) { c_abi::block_set_gap_open_C_aaprofile(p_, i, gap); }
  void set_gap_close_C(size_t i, int8_t gap) { c_abi::block_set_gap_close_C_aaprofile(p_, i, gap); }
  void set_gap_open_R(size_t i, int8_t gap) { c_abi::block_set_gap_open_R_aaprofile(p_, i, gap); }
  void set_all_gap_open_C(int8_t gap) { c_abi::block_set_all_gap_open_C_aaprofile(p_, gap); }
  void set_all_gap_close_C(int8_t gap) { c_abi::block_set_all_gap_close_C_aaprofile(p_, gap); }
  void set_all_gap_open_R(int8_t gap) { c_abi::block_set_all_gap_open_R_aaprofile(p_, gap); }
  int8_t get(size_t i, uint8_t b) const { return c_abi::block_get_aaprofile(p_, i, b); }
  int8_t get_gap_extend() const { return c_abi::block_get_gap_extend_aaprofile(p_); }
  const c_abi::AAProfile* handle() const { return p_; }
 private:
  c_abi::AAProfile* p_;
};

// ---- Cigar (cigar.rs:42-163) ---------------------------------------------------------------------------------
class Cigar {
 public:
  Cigar(size_t query_len, size_t reference_len) { s_.reserve(query_len + reference_len + 5); }   // cigar.rs:50
  void clear(size_t, size_t) { s_.clear(); }
  size_t len() const { return s_.size(); }                                          // cigar.rs:96-98 (without the sentinel)
  OpLen get(size_t i) const { return s_.at(i); }                                    // cigar.rs:104-106, forward order
  std::vector<OpLen> to_vec() const { return s_; }
  std::string to_string() const {                                                   // cigar.rs:147-163
    static const char ops[6] = {'?', 'M', '=', 'X', 'I', 'D'};
    std::string out;
    for (const OpLen& o : s_) { out += std::to_string(o.len); out += ops[(uint8_t)o.op <= 5 ? (uint8_t)o.op : 0]; }
    return out;
  }
  void assign_runs(const uint32_t* runs, size_t n) {
    s_.clear();
    for (size_t k = 0; k < n; k++) s_.push_back(OpLen{(Operation)(runs[k] & 15u), (size_t)(runs[k] >> 4)});
  }
 private:
  std::vector<OpLen> s_;
};

// ---- Block (scan_block.rs:787-1241) -------------------------------------------------------------------------
template <bool TRACE, bool X_DROP, bool LOCAL_START = false, bool FREE_QUERY_START_GAPS = false, bool FREE_QUERY_END_GAPS = false>
class Block {
 public:
  // Block::new (scan_block.rs:798-805)
  Block(size_t query_len, size_t reference_len, size_t max_size) : query_len_(query_len), reference_len_(reference_len), max_size_(max_size) {
    detail::require(max_size && !(max_size & (max_size - 1)), "Block size must be a power of two!");
  }
  Block(const Block&) = delete;
  Block& operator=(const Block&) = delete;
  ~Block() { if (last_) c_abi::ba_batch_free(last_); }

  // Block::align (scan_block.rs:847-878)
  template <class M>
  void align(const PaddedBytes& query, const PaddedBytes& reference, const M& matrix, Gaps gaps, SizeRange size, int32_t x_drop) {
    c_abi::BaConfig cfg = config(M::kScoring, matrix.data(), gaps, size, x_drop);
    check_fit(query.len(), reference.len(), size);
    const uint64_t qo[2] = {0, query.len()}, ro[2] = {0, reference.len()};
    reset();
    detail::check(c_abi::ba_batch_upload(detail::aligner(), &cfg, 1, query.raw().data(), qo, reference.raw().data(), ro, &last_));
    finish(query.len(), reference.len());
  }
  // Block::align_profile (scan_block.rs:942-968)
  void align_profile(const PaddedBytes& query, const AAProfile& profile, SizeRange size, int32_t x_drop) {
    c_abi::BaConfig cfg = config(c_abi::BA_SCORING_PROFILE, nullptr, Gaps{0, 0}, size, x_drop);
    check_fit(query.len(), profile.len(), size);
    const uint64_t qo[2] = {0, query.len()};
    const c_abi::AAProfile* pp[1] = {profile.handle()};
    reset();
    detail::check(c_abi::ba_batch_upload_profiles(detail::aligner(), &cfg, 1, query.raw().data(), qo, pp, &last_));
    finish(query.len(), profile.len());
  }
  // Block::align_exp (scan_block.rs:884-902): min size doubles until the score reaches target_score
  template <class M>
  std::optional<size_t> align_exp(const PaddedBytes& query, const PaddedBytes& reference, const M& matrix, Gaps gaps, SizeRange size,
                                  int32_t x_drop, int32_t target_score) {
    for (size_t s = size.min; s <= size.max; s *= 2) {
      align(query, reference, matrix, gaps, SizeRange{s, size.max}, x_drop);
      if (res_.score >= target_score) return s;
    }
    return std::nullopt;
  }
  // Block::align_profile_exp (scan_block.rs:974-992)
  std::optional<size_t> align_profile_exp(const PaddedBytes& query, const AAProfile& profile, SizeRange size, int32_t x_drop, int32_t target_score) {
    for (size_t s = size.min; s <= size.max; s *= 2) {
      align_profile(query, profile, SizeRange{s, size.max}, x_drop);
      if (res_.score >= target_score) return s;
    }
    return std::nullopt;
  }
  AlignResult res() const { return res_; }                                           // scan_block.rs:1235

  // Trace (scan_block.rs:1344-1692): only the traceback side is visible to users
  class Trace {
   public:
    void cigar(size_t query_idx, size_t reference_idx, Cigar& cigar) const { b_->walk(query_idx, reference_idx, false, cigar); }   // :1469-1473
    void cigar_eq(const PaddedBytes&, const PaddedBytes&, size_t query_idx, size_t reference_idx, Cigar& cigar) const {           // :1475-1480
      b_->walk(query_idx, reference_idx, true, cigar);
    }
   private:
    friend class Block;
    explicit Trace(const Block* b) : b_(b) {}
    const Block* b_;
  };
  Trace trace() const {                                                              // scan_block.rs:1241
    detail::require(TRACE, "assertion failed: TRACE");
    return Trace(this);
  }

 private:
  static constexpr int kFlags = (TRACE ? c_abi::BA_TRACE : 0) | (X_DROP ? c_abi::BA_XDROP : 0) | (LOCAL_START ? c_abi::BA_LOCAL_START : 0) |
                                (FREE_QUERY_START_GAPS ? c_abi::BA_FREE_QUERY_START_GAPS : 0) | (FREE_QUERY_END_GAPS ? c_abi::BA_FREE_QUERY_END_GAPS : 0);
  c_abi::BaConfig config(int scoring, const void* matrix, Gaps gaps, SizeRange size, int32_t x_drop) const {
    c_abi::BaConfig cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.scoring = scoring; cfg.flags = kFlags; cfg.matrix = matrix; cfg.gaps = gaps; cfg.size = size; cfg.x_drop = x_drop;
    return cfg;
  }
  void check_fit(size_t ql, size_t rl, SizeRange size) const {                       // Allocated::clear, scan_block.rs:1324-1326
    detail::require(ql + rl <= query_len_ + reference_len_, "Block: sequences are longer than the lengths the block was created with");
    detail::require((size.max < 16 ? 16 : size.max) <= max_size_, "Block: max block size exceeds the size the block was created with");
  }
  void reset() { if (last_) { c_abi::ba_batch_free(last_); last_ = nullptr; } }
  void finish(size_t ql, size_t rl) {
    detail::check(c_abi::ba_batch_run(last_, nullptr));
    detail::check(c_abi::ba_batch_download(last_, &res_));
    last_qlen_ = ql; last_rlen_ = rl;
  }
  void walk(size_t qi, size_t rj, bool eq, Cigar& cigar) const {
    detail::require(last_ != nullptr, "Block: no alignment to trace back");
    detail::require(qi <= last_qlen_ && rj <= last_rlen_, "Traceback cigar end position must be in bounds!");   // scan_block.rs:1483
    const uint32_t* runs = nullptr; size_t n = 0;
    detail::check(c_abi::ba_batch_traceback(last_, 0, qi, rj, eq ? 1 : 0, &runs, &n));
    cigar.assign_runs(runs, n);
  }
  size_t query_len_, reference_len_, max_size_;
  AlignResult res_{0, 0, 0};
  c_abi::BaBatch* last_ = nullptr;
  size_t last_qlen_ = 0, last_rlen_ = 0;
};

// ---- batch extension (not in the reference): many pairs in one call -----------------------------------------
// -> results in input order; with TRACE also the CIGAR strings (cigar_eq selects '='/'X' instead of 'M').
template <bool TRACE, bool X_DROP, class M>
std::vector<AlignResult> align_batch(const std::vector<std::string_view>& queries, const std::vector<std::string_view>& references,
                                     const M& matrix, Gaps gaps, SizeRange size, int32_t x_drop,
                                     std::vector<std::string>* cigars = nullptr, bool cigar_eq = true) {
  detail::require(queries.size() == references.size(), "align_batch: queries and references differ in number");
  const size_t n = queries.size();
  std::vector<uint64_t> qo(n + 1, 0), ro(n + 1, 0);
  for (size_t k = 0; k < n; k++) { qo[k + 1] = qo[k] + queries[k].size(); ro[k + 1] = ro[k] + references[k].size(); }
  std::vector<uint8_t> qa(qo[n] + 1), ra(ro[n] + 1);
  for (size_t k = 0; k < n; k++) {
    std::memcpy(qa.data() + qo[k], queries[k].data(), queries[k].size());
    std::memcpy(ra.data() + ro[k], references[k].data(), references[k].size());
  }
  c_abi::BaConfig cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.scoring = M::kScoring; cfg.flags = (TRACE ? c_abi::BA_TRACE : 0) | (X_DROP ? c_abi::BA_XDROP : 0); cfg.matrix = matrix.data();
  cfg.gaps = gaps; cfg.size = size; cfg.x_drop = x_drop; cfg.cigar_eq = cigar_eq ? 1 : 0;
  std::vector<AlignResult> out(n);
  c_abi::BaBatch* b = nullptr;
  detail::check(c_abi::ba_batch_upload(detail::aligner(), &cfg, n, qa.data(), qo.data(), ra.data(), ro.data(), &b));
  int rc = c_abi::ba_batch_run(b, nullptr);
  if (!rc) rc = c_abi::ba_batch_download(b, out.data());
  if (!rc && TRACE && cigars) {
    cigars->assign(n, std::string());
    for (size_t k = 0; k < n && !rc; k++) {
      const uint32_t* runs = nullptr; size_t nr = 0;
      rc = c_abi::ba_batch_cigar(b, k, &runs, &nr);
      if (!rc) {
        Cigar c(0, 0);
        c.assign_runs(runs, nr);
        (*cigars)[k] = c.to_string();
      }
    }
  }
  c_abi::ba_batch_free(b);
  detail::check(rc);
  return out;
}

}  // namespace block_aligner
