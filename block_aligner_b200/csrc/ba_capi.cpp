// ba_capi.cpp -- Part 1 of include/block_aligner_b200.h: the reference's own C API
// (reference: c/block_aligner.h, src/ffi.rs), plus the small host helpers of Part 2.
//
// Scoring tables, padded strings, profiles and CIGAR containers are plain host objects. Every
// block_align_* call marshals its single pair through the batch path and runs on the GPU; there is
// no CPU implementation of the aligner in this library.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <string>

#include "../../include/block_aligner_b200.h"
#include "ba_host.h"
#include "matrices_data.h"

using namespace ba;

[[noreturn]] static void die(const char* msg) {
  // the reference panics with this message and the release profile aborts (Cargo.toml:41)
  fprintf(stderr, "block-aligner-b200: %s\n", msg);
  abort();
}
#define REQUIRE(cond, msg) do { if (!(cond)) die(msg); } while (0)

// ---- statics (scores.rs:277-311) -----------------------------------------------------------------
#define NUC_IDX(a, b) ((((a)&7) * 16) + ((b)&15))
static constexpr NucMatrix make_nuc(int8_t match, int8_t mismatch) {   // scores.rs:150-164
  NucMatrix m{};
  for (int k = 0; k < 128; k++) m.scores[k] = INT8_MIN;
  const char alpha[5] = {'A', 'T', 'C', 'G', 'N'};
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) m.scores[NUC_IDX(alpha[i], alpha[j])] = (i == j) ? match : mismatch;
  return m;
}
extern "C" {
extern const NucMatrix NW1 = make_nuc(1, -1);
extern const AAMatrix BLOSUM45 = {{BA_BLOSUM45_VALUES}};
extern const AAMatrix BLOSUM50 = {{BA_BLOSUM50_VALUES}};
extern const AAMatrix BLOSUM62 = {{BA_BLOSUM62_VALUES}};
extern const AAMatrix BLOSUM80 = {{BA_BLOSUM80_VALUES}};
extern const AAMatrix BLOSUM90 = {{BA_BLOSUM90_VALUES}};
extern const AAMatrix PAM100 = {{BA_PAM100_VALUES}};
extern const AAMatrix PAM120 = {{BA_PAM120_VALUES}};
extern const AAMatrix PAM160 = {{BA_PAM160_VALUES}};
extern const AAMatrix PAM200 = {{BA_PAM200_VALUES}};
extern const AAMatrix PAM250 = {{BA_PAM250_VALUES}};
extern const ByteMatrix BYTES1 = {1, -1};
}

// ---- host helpers used by the runtime -------------------------------------------------------------
namespace ba { namespace host {
size_t profile_len(const AAProfile* p) { return p->str_len; }
size_t profile_curr_len(const AAProfile* p) { return p->curr_len; }
int profile_gap_extend(const AAProfile* p) { return p->gap_extend; }
size_t profile_used_len(const AAProfile* p) { return std::min(p->hi_pos + 1, p->curr_len); }
void profile_export(const AAProfile* p, size_t np, int8_t* pos_aa, int16_t* oc, int16_t* cc, int16_t* orr) {
  memcpy(pos_aa, p->pos_aa.data(), np * 32);
  memcpy(oc, p->gap_open_C.data(), np * 2);
  memcpy(cc, p->gap_close_C.data(), np * 2);
  memcpy(orr, p->gap_open_R.data(), np * 2);
}
}}

extern "C" {

// ---- AAMatrix (ffi.rs:27-48; scores.rs:48-61, 89-98) ------------------------------------------------
AAMatrix* block_new_simple_aamatrix(int8_t match_score, int8_t mismatch_score) {
  AAMatrix* m = new AAMatrix();
  for (int k = 0; k < 27 * 32; k++) m->scores[k] = INT8_MIN;
  for (int i = 0; i < 26; i++)
    for (int j = 0; j < 26; j++) m->scores[i * 32 + j] = (i == j) ? match_score : mismatch_score;
  return m;
}
void block_set_aamatrix(AAMatrix* m, uint8_t a, uint8_t b, int8_t score) {
  a = host::upper(a); b = host::upper(b);
  REQUIRE(a >= 'A' && a <= 'Z' + 1 && b >= 'A' && b <= 'Z' + 1, "AAMatrix::set: byte out of range");
  m->scores[(a - 'A') * 32 + (b - 'A')] = score;
  m->scores[(b - 'A') * 32 + (a - 'A')] = score;
}
void block_free_aamatrix(AAMatrix* m) { delete m; }

// ---- NucMatrix helpers (the reference's C header has none; Rust API scores.rs:150-192) ------------
NucMatrix* ba_new_simple_nucmatrix(int8_t match_score, int8_t mismatch_score) {
  NucMatrix* m = new NucMatrix();
  *m = make_nuc(match_score, mismatch_score);
  return m;
}
void ba_set_nucmatrix(NucMatrix* m, uint8_t a, uint8_t b, int8_t score) {
  a = host::upper(a); b = host::upper(b);
  REQUIRE(a >= 'A' && a <= 'Z' && b >= 'A' && b <= 'Z', "NucMatrix::set: byte out of range");
  m->scores[NUC_IDX(a, b)] = score;
  m->scores[NUC_IDX(b, a)] = score;
}
void ba_free_nucmatrix(NucMatrix* m) { delete m; }

// ---- AAProfile (ffi.rs:53-190; scores.rs:470-715) ---------------------------------------------------
AAProfile* block_new_aaprofile(uintptr_t str_len, uintptr_t block_size, int8_t gap_extend) {
  AAProfile* p = new AAProfile();
  p->max_len = str_len + block_size + 1;
  p->pos_aa.assign(p->max_len * 32, INT8_MIN);
  p->gap_open_C.assign(p->max_len, (int16_t)INT8_MIN);
  p->gap_close_C.assign(p->max_len, (int16_t)INT8_MIN);
  p->gap_open_R.assign(p->max_len, (int16_t)INT8_MIN);
  p->gap_extend = gap_extend;
  p->curr_len = p->max_len;
  p->str_len = str_len;
  p->hi_pos = 0;
  return p;
}
uintptr_t block_len_aaprofile(const AAProfile* p) { return p->str_len; }
void block_clear_aaprofile(AAProfile* p, uintptr_t str_len, uintptr_t block_size) {
  const size_t cl = str_len + block_size + 1;
  REQUIRE(cl <= p->max_len, "AAProfile::clear: length exceeds the allocated profile");
  std::fill(p->pos_aa.begin(), p->pos_aa.begin() + cl * 32, INT8_MIN);
  std::fill(p->gap_open_C.begin(), p->gap_open_C.begin() + cl, (int16_t)INT8_MIN);
  std::fill(p->gap_close_C.begin(), p->gap_close_C.begin() + cl, (int16_t)INT8_MIN);
  std::fill(p->gap_open_R.begin(), p->gap_open_R.begin() + cl, (int16_t)INT8_MIN);
  p->str_len = str_len;
  p->curr_len = cl;
  p->hi_pos = 0;
}
void block_set_aaprofile(AAProfile* p, uintptr_t i, uint8_t b, int8_t score) {
  b = host::upper(b);
  REQUIRE(b >= 'A' && b <= 'Z' + 1, "AAProfile::set: byte out of range");
  REQUIRE(i < p->curr_len, "AAProfile::set: position out of range");
  p->pos_aa[i * 32 + (b - 'A')] = score;
  p->hi_pos = std::max(p->hi_pos, (size_t)i);
}
static void set_all_core(AAProfile* p, const uint8_t* order, size_t order_len, const int8_t* scores, size_t scores_len,
                         size_t ls, size_t rs, bool rev) {   // scores.rs:677-714
  uint8_t o[32];
  for (int k = 0; k < 32; k++) o[k] = 26;
  REQUIRE(order_len <= 32 && order_len > 0, "AAProfile::set_all: order too long");
  for (size_t k = 0; k < order_len; k++) {
    const uint8_t b = host::upper(order[k]);
    REQUIRE(b >= 'A' && b <= 'Z' + 1, "AAProfile::set_all: byte out of range");
    o[k] = (uint8_t)(b - 'A');
  }
  REQUIRE(scores_len / order_len == p->str_len, "AAProfile::set_all: scores length does not match the profile length");
  size_t si = 0;
  for (size_t n = 0; n < p->str_len; n++) {
    const size_t i = rev ? p->str_len - n : 1 + n;
    for (size_t j = 0; j < order_len; j++, si++)
      p->pos_aa[i * 32 + o[j]] = (int8_t)((int8_t)(scores[si] << ls) >> rs);
  }
  p->hi_pos = std::max(p->hi_pos, p->str_len);
}
void block_set_all_aaprofile(AAProfile* p, const uint8_t* order, uintptr_t order_len, const int8_t* scores,
                             uintptr_t scores_len, uintptr_t left_shift, uintptr_t right_shift) {
  set_all_core(p, order, order_len, scores, scores_len, left_shift, right_shift, false);
}
void block_set_all_rev_aaprofile(AAProfile* p, const uint8_t* order, uintptr_t order_len, const int8_t* scores,
                                 uintptr_t scores_len, uintptr_t left_shift, uintptr_t right_shift) {
  set_all_core(p, order, order_len, scores, scores_len, left_shift, right_shift, true);
}
void block_set_gap_open_C_aaprofile(AAProfile* p, uintptr_t i, int8_t gap) {
  REQUIRE(gap < 0, "Gap open cost must be negative!");
  REQUIRE(i < p->curr_len, "AAProfile: position out of range");
  p->gap_open_C[i] = gap;
  p->hi_pos = std::max(p->hi_pos, (size_t)i);
}
void block_set_gap_close_C_aaprofile(AAProfile* p, uintptr_t i, int8_t gap) {
  REQUIRE(i < p->curr_len, "AAProfile: position out of range");
  p->gap_close_C[i] = gap;
  p->hi_pos = std::max(p->hi_pos, (size_t)i);
}
void block_set_gap_open_R_aaprofile(AAProfile* p, uintptr_t i, int8_t gap) {
  REQUIRE(gap < 0, "Gap open cost must be negative!");
  REQUIRE(i < p->curr_len, "AAProfile: position out of range");
  p->gap_open_R[i] = gap;
  p->hi_pos = std::max(p->hi_pos, (size_t)i);
}
void block_set_all_gap_open_C_aaprofile(AAProfile* p, int8_t gap) {
  REQUIRE(gap < 0, "Gap open cost must be negative!");
  std::fill(p->gap_open_C.begin(), p->gap_open_C.begin() + p->str_len + 1, (int16_t)gap);
  p->hi_pos = std::max(p->hi_pos, p->str_len);
}
void block_set_all_gap_close_C_aaprofile(AAProfile* p, int8_t gap) {
  std::fill(p->gap_close_C.begin(), p->gap_close_C.begin() + p->str_len + 1, (int16_t)gap);
  p->hi_pos = std::max(p->hi_pos, p->str_len);
}
void block_set_all_gap_open_R_aaprofile(AAProfile* p, int8_t gap) {
  REQUIRE(gap < 0, "Gap open cost must be negative!");
  std::fill(p->gap_open_R.begin(), p->gap_open_R.begin() + p->str_len + 1, (int16_t)gap);
  p->hi_pos = std::max(p->hi_pos, p->str_len);
}
int8_t block_get_aaprofile(const AAProfile* p, uintptr_t i, uint8_t b) {
  b = host::upper(b);
  REQUIRE(b >= 'A' && b <= 'Z' + 1, "AAProfile::get: byte out of range");
  return p->pos_aa[i * 32 + (b - 'A')];
}
int8_t block_get_gap_extend_aaprofile(const AAProfile* p) { return p->gap_extend; }
void block_free_aaprofile(AAProfile* p) { delete p; }

// ---- Cigar (ffi.rs:195-224; cigar.rs:47-94) -----------------------------------------------------------
Cigar* block_new_cigar(uintptr_t query_len, uintptr_t reference_len) {
  Cigar* c = new Cigar();
  c->s.assign(query_len + reference_len + 5, CigarRun{0, 0});
  c->idx = 1;
  return c;
}
OpLen block_get_cigar(const Cigar* c, uintptr_t i) {
  REQUIRE(i < c->idx - 1, "Cigar::get: index out of bounds");
  const CigarRun& r = c->s[c->idx - 1 - i];
  OpLen o; o.op = (Operation)r.op; o.len = r.len;
  return o;
}
uintptr_t block_len_cigar(const Cigar* c) { return c->idx - 1; }
void block_free_cigar(Cigar* c) { delete c; }

// ---- PaddedBytes (ffi.rs:229-257; scan_block.rs:1798-1822) ------------------------------------------
PaddedBytes* block_new_padded_aa(uintptr_t len, uintptr_t max_size) {
  PaddedBytes* p = new PaddedBytes();
  p->s.assign(1 + len + max_size, host::null_code(kAA));
  p->len = len;
  return p;
}
static void set_bytes(PaddedBytes* p, const uint8_t* s, size_t len, size_t block_size, bool rev) {
  REQUIRE(1 + len + block_size <= p->s.size(), "PaddedBytes::set_bytes: string longer than the allocated capacity");
  const uint8_t nul = host::null_code(kAA);
  p->s[0] = nul;
  for (size_t k = 0; k < len; k++) {
    bool ok;
    p->s[1 + k] = host::convert_char(kAA, s[rev ? len - 1 - k : k], &ok);
    REQUIRE(ok, "AAMatrix::convert_char: byte out of range");
  }
  for (size_t k = 0; k < block_size; k++) p->s[1 + len + k] = nul;
  p->len = len;
}
void block_set_bytes_padded_aa(PaddedBytes* p, const uint8_t* s, uintptr_t len, uintptr_t max_size) { set_bytes(p, s, len, max_size, false); }
void block_set_bytes_rev_padded_aa(PaddedBytes* p, const uint8_t* s, uintptr_t len, uintptr_t max_size) { set_bytes(p, s, len, max_size, true); }
void block_free_padded_aa(PaddedBytes* p) { delete p; }

// ---- BA_INPUT_NUC4 host helper: ASCII -> BAM nibble codes ("=ACMGRSVTWYHKDBN") ---------------------------------
size_t ba_pack_nuc4(const uint8_t* ascii, size_t len, uint8_t* packed, uint64_t nibble_off) {
  static const char code[] = "=ACMGRSVTWYHKDBN";
  uint8_t lut[256];
  memset(lut, 0, sizeof(lut));
  for (int c = 1; c < 16; c++) { lut[(uint8_t)code[c]] = (uint8_t)c; lut[(uint8_t)(code[c] + 32)] = (uint8_t)c; }
  for (size_t k = 0; k < len; k++) {
    const uint8_t v = lut[ascii[k]];
    if (!v) return k + 1;
    const uint64_t nb = nibble_off + k;
    uint8_t& b = packed[nb >> 1];
    b = (nb & 1) ? (uint8_t)((b & 0xf0u) | v) : (uint8_t)((b & 0x0fu) | (v << 4));
  }
  return 0;
}

// ---- lib.rs:109-111 --------------------------------------------------------------------------------
uintptr_t ba_percent_len(uintptr_t len, float p) {
  // Rust's f32::round rounds half away from zero, like roundf
  uint64_t v = (uint64_t)roundf(p * (float)len);
  if (v < 32) v = 32;
  uint64_t n = 1;
  while (n < v) n <<= 1;
  if (n > (1u << 14)) n = 1u << 14;
  return (uintptr_t)n;
}

// ---- cigar.rs:147-163 ------------------------------------------------------------------------------
size_t ba_cigar_format(const uint32_t* runs, size_t n_runs, char* out, size_t cap) {
  static const char ops[6] = {'?', 'M', '=', 'X', 'I', 'D'};
  size_t w = 0;
  for (size_t k = 0; k < n_runs; k++) {
    const uint32_t op = runs[k] & 15u;
    if (op == 0 || op > 5) continue;
    char tmp[24];
    const int m = snprintf(tmp, sizeof(tmp), "%u%c", runs[k] >> 4, ops[op]);
    for (int t = 0; t < m; t++) { if (out && w + 1 < cap) out[w] = tmp[t]; w++; }
  }
  if (out && cap) out[w < cap ? w : cap - 1] = 0;
  return w;
}

// ---- Block (ffi.rs:262-403) ------------------------------------------------------------------------
struct BlockObj {
  int flags;
  size_t query_len, reference_len, max_size;
  AlignResult res;
  BaBatch* last;     // kept so that block_cigar_* can walk the trace that is still on the device
  size_t last_qlen, last_rlen;
};

static BaAligner* g_aligner = nullptr;
static std::mutex g_mu;
static BaAligner* the_aligner() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_aligner) {
    int dev = 0;
    if (const char* e = getenv("BA_DEVICE")) dev = atoi(e);
    const int rc = ba_create(dev, &g_aligner);
    if (rc) { fprintf(stderr, "block-aligner-b200: %s (%s)\n", ba_error_string(rc), ba_last_error_message()); abort(); }
  }
  return g_aligner;
}

int ba_batch_traceback(BaBatch* b, size_t k, size_t query_idx, size_t reference_idx, int eq,
                       const uint32_t** runs, size_t* n_runs);  // ba_runtime.cu

static BlockHandle block_new(uintptr_t ql, uintptr_t rl, uintptr_t max_size, int flags) {
  REQUIRE(max_size && !(max_size & (max_size - 1)), "Block size must be a power of two!");   // scan_block.rs:799
  BlockObj* o = new BlockObj();
  o->flags = flags; o->query_len = ql; o->reference_len = rl; o->max_size = max_size;
  o->res = AlignResult{0, 0, 0}; o->last = nullptr; o->last_qlen = o->last_rlen = 0;
  return (BlockHandle)o;
}
static void block_free(BlockHandle h) {
  BlockObj* o = (BlockObj*)h;
  if (!o) return;
  if (o->last) ba_batch_free(o->last);
  delete o;
}
static void check_rc(int rc) {
  if (rc == BA_OK) return;
  fprintf(stderr, "block-aligner-b200: %s (%s)\n", ba_error_string(rc), ba_last_error_message());
  abort();
}
// PaddedBytes keeps converted codes; the batch path converts raw bytes on the device, so undo it.
static std::vector<uint8_t> raw_of(const PaddedBytes* p) {
  std::vector<uint8_t> v(p->len);
  for (size_t k = 0; k < p->len; k++) v[k] = (uint8_t)(p->s[1 + k] + 'A');
  return v;
}
static void run_one(BlockObj* o, const PaddedBytes* q, const PaddedBytes* r, const AAProfile* prof,
                    const AAMatrix* m, Gaps g, SizeRange s, int32_t x) {
  const size_t rlen = prof ? prof->str_len : r->len;
  const size_t mx = s.max < (size_t)kL ? kL : s.max;
  // Allocated::clear (scan_block.rs:1324-1326)
  REQUIRE(q->len + rlen <= o->query_len + o->reference_len, "Block: sequences are longer than the lengths the block was created with");
  REQUIRE(mx <= o->max_size, "Block: max block size exceeds the size the block was created with");
  BaConfig cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.scoring = prof ? BA_SCORING_PROFILE : BA_SCORING_AA;
  cfg.flags = o->flags; cfg.matrix = prof ? nullptr : (const void*)m->scores;
  cfg.gaps = g; cfg.size = s; cfg.x_drop = x; cfg.cigar_eq = 0;
  std::vector<uint8_t> qb = raw_of(q), rb;
  const uint64_t qo[2] = {0, qb.size()};
  uint64_t ro[2] = {0, 0};
  if (o->last) { ba_batch_free(o->last); o->last = nullptr; }
  BaBatch* b = nullptr;
  int rc;
  if (prof) {
    const AAProfile* pp[1] = {prof};
    rc = ba_batch_upload_profiles(the_aligner(), &cfg, 1, qb.data(), qo, pp, &b);
  } else {
    rb = raw_of(r); ro[1] = rb.size();
    rc = ba_batch_upload(the_aligner(), &cfg, 1, qb.data(), qo, rb.data(), ro, &b);
  }
  check_rc(rc);
  check_rc(ba_batch_run(b, nullptr));
  check_rc(ba_batch_download(b, &o->res));
  o->last = b; o->last_qlen = q->len; o->last_rlen = rlen;
}
static void cigar_one(BlockObj* o, uintptr_t qi, uintptr_t rj, bool eq, const PaddedBytes* q, const PaddedBytes* r, Cigar* c) {
  REQUIRE(o->flags & BA_TRACE, "assertion failed: TRACE");                                   // scan_block.rs:1242
  REQUIRE(o->last, "Block: no alignment to trace back");
  REQUIRE(qi <= o->last_qlen && rj <= o->last_rlen, "Traceback cigar end position must be in bounds!");  // :1483
  if (eq) REQUIRE(q && r, "cigar_eq needs both sequences");
  const uint32_t* runs = nullptr; size_t n = 0;
  check_rc(ba_batch_traceback(o->last, 0, qi, rj, eq ? 1 : 0, &runs, &n));
  // Cigar::clear + add, stored reversed behind a sentinel (cigar.rs:58-79)
  REQUIRE(qi + rj + 5 <= c->s.size() || n + 1 <= c->s.size(), "Cigar: capacity too small");
  std::fill(c->s.begin(), c->s.end(), CigarRun{0, 0});
  c->idx = 1;
  for (size_t k = 0; k < n; k++) {
    const uint32_t w = runs[n - 1 - k];
    c->s[c->idx].op = (uint8_t)(w & 15u); c->s[c->idx].len = w >> 4;
    c->idx++;
  }
}

#define BA_DEFINE_BLOCK_API(SUFFIX, CIGAR, CIGAR_EQ, FLAGS)                                                          \
  BlockHandle block_new_##SUFFIX(uintptr_t ql, uintptr_t rl, uintptr_t ms) { return block_new(ql, rl, ms, FLAGS); }   \
  void block_align_##SUFFIX(BlockHandle b, const PaddedBytes* q, const PaddedBytes* r, const AAMatrix* m, Gaps g,      \
                            SizeRange s, int32_t x) { run_one((BlockObj*)b, q, r, nullptr, m, g, s, x); }              \
  void block_align_profile_##SUFFIX(BlockHandle b, const PaddedBytes* q, const AAProfile* r, SizeRange s, int32_t x) { \
    Gaps g = {0, 0}; run_one((BlockObj*)b, q, nullptr, r, nullptr, g, s, x); }                                        \
  AlignResult block_res_##SUFFIX(BlockHandle b) { return ((BlockObj*)b)->res; }                                        \
  void CIGAR(BlockHandle b, uintptr_t qi, uintptr_t rj, Cigar* c) { cigar_one((BlockObj*)b, qi, rj, false, nullptr, nullptr, c); } \
  void CIGAR_EQ(BlockHandle b, const PaddedBytes* q, const PaddedBytes* r, uintptr_t qi, uintptr_t rj, Cigar* c) {    \
    cigar_one((BlockObj*)b, qi, rj, true, q, r, c); }                                                                 \
  void block_free_##SUFFIX(BlockHandle b) { block_free(b); }

BA_DEFINE_BLOCK_API(aa, _block_cigar_aa, _block_cigar_eq_aa, 0)
BA_DEFINE_BLOCK_API(aa_xdrop, _block_cigar_aa_xdrop, _block_cigar_eq_aa_xdrop, BA_XDROP)
BA_DEFINE_BLOCK_API(aa_trace, block_cigar_aa_trace, block_cigar_eq_aa_trace, BA_TRACE)
BA_DEFINE_BLOCK_API(aa_trace_xdrop, block_cigar_aa_trace_xdrop, block_cigar_eq_aa_trace_xdrop, BA_TRACE | BA_XDROP)

}  // extern "C"
