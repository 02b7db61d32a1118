// The reference's own unit tests (src/scan_block.rs:1908-2230, src/lib.rs:8-35) through the C++ mirror of the Rust
// surface (include/block_aligner_b200.hpp). Known answers are the reference's; nothing here computes an expectation.
// Built by tests/test_cpp_api.py against the emulated library (CPU) and against the product library (GPU).
#include <cstdio>
#include <string>

#include "block_aligner_b200.hpp"

using namespace block_aligner;

static int g_checks = 0, g_failed = 0;
#define CHECK(cond) do { g_checks++; if (!(cond)) { g_failed++; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)

template <class M, class B, class MAT>
static AlignResult run(B& a, const char* q, const char* r, const MAT& m, Gaps g, size_t lo, size_t hi, int x, size_t pad) {
  auto rp = PaddedBytes::from_bytes<M>(r, pad);
  auto qp = PaddedBytes::from_bytes<M>(q, pad);
  a.align(qp, rp, m, g, SizeRange{lo, hi}, x);
  return a.res();
}

static void test_no_x_drop() {          // scan_block.rs:1908-1992
  const Gaps g{-11, -1};
  Block<false, false> a(100, 100, 16);
  struct { const char *q, *r; int score; } aa[] = {{"", "", 0}, {"", "AAAA", -14}, {"AAAA", "", -14}, {"AARA", "AAAA", 11},
      {"AARAAAA", "AAAAAAAA", 12}, {"AAAA", "AAAA", 16}, {"RRRR", "AAAA", -4}, {"AAA", "AAAA", 1}};
  for (auto& t : aa) CHECK(run<AAMatrix>(a, t.q, t.r, BLOSUM62(), g, 16, 16, 0, 16).score == t.score);
  const Gaps g2{-2, -1};
  const std::string a32(32, 'A'), t32(32, 'T');
  std::string ta32; for (int i = 0; i < 16; i++) ta32 += "TA";
  CHECK(run<NucMatrix>(a, "ATAA", "AAAN", NW1(), g2, 16, 16, 0, 16).score == 0);
  CHECK(run<NucMatrix>(a, a32.c_str(), a32.c_str(), NW1(), g2, 16, 16, 0, 16).score == 32);
  CHECK(run<NucMatrix>(a, t32.c_str(), a32.c_str(), NW1(), g2, 16, 16, 0, 16).score == -32);
  CHECK(run<NucMatrix>(a, ta32.c_str(), a32.c_str(), NW1(), g2, 16, 16, 0, 16).score == 0);
  CHECK(run<NucMatrix>(a, "TTTTTTTTAAAAAAATTTTTTTTT", "TTAAAAAAATTTTTTTTTTTT", NW1(), g2, 16, 16, 0, 16).score == 7);
  CHECK(run<NucMatrix>(a, "C", "AAAA", NW1(), g2, 16, 16, 0, 16).score == -5);
  CHECK(run<NucMatrix>(a, "AAAA", "C", NW1(), g2, 16, 16, 0, 16).score == -5);
}

static void test_x_drop() {             // scan_block.rs:1994-2050
  const Gaps g{-11, -1};
  Block<false, true> a(100, 100, 16);
  CHECK((run<AAMatrix>(a, "", "", BLOSUM62(), g, 16, 16, 1, 16) == AlignResult{0, 0, 0}));
  CHECK((run<AAMatrix>(a, "", "AAAA", BLOSUM62(), g, 16, 16, 1, 16) == AlignResult{0, 0, 0}));
  CHECK((run<AAMatrix>(a, "AAAA", "", BLOSUM62(), g, 16, 16, 1, 16) == AlignResult{0, 0, 0}));
  CHECK((run<AAMatrix>(a, "AAAAAA", "AAARRA", BLOSUM62(), g, 16, 16, 1, 16) == AlignResult{14, 6, 6}));
  const std::string q44(44, 'A'), r44 = std::string(15, 'A') + std::string(16, 'R') + std::string(13, 'A');
  CHECK((run<AAMatrix>(a, q44.c_str(), r44.c_str(), BLOSUM62(), g, 16, 16, 1, 16) == AlignResult{60, 15, 15}));
  Block<true, true> big(2048, 2048, 2048);
  const std::string l(2048, 'A');
  CHECK((run<AAMatrix>(big, l.c_str(), l.c_str(), BLOSUM62(), g, 2048, 2048, 100, 2048) == AlignResult{8192, 2048, 2048}));
  Block<true, true> zero(0, 0, 16);
  CHECK((run<AAMatrix>(zero, "", "", BLOSUM62(), g, 16, 16, 1, 16) == AlignResult{0, 0, 0}));
  Block<true, true> four(4, 4, 16);
  CHECK((run<AAMatrix>(four, "", "AAAA", BLOSUM62(), g, 16, 16, 1, 16) == AlignResult{0, 0, 0}));
  CHECK((run<AAMatrix>(four, "AAAA", "", BLOSUM62(), g, 16, 16, 1, 16) == AlignResult{0, 0, 0}));
}

static void test_trace() {              // scan_block.rs:2052-2103
  const Gaps g{-11, -1}, g2{-2, -1}, g3{-5, -2};
  Cigar cigar(100, 100);
  Block<true, false> a(100, 100, 16);
  {
    auto r = PaddedBytes::from_bytes<AAMatrix>("AAARRA", 16), q = PaddedBytes::from_bytes<AAMatrix>("AAAAAA", 16);
    a.align(q, r, BLOSUM62(), g, {16, 16}, 0);
    const AlignResult res = a.res();
    CHECK((res == AlignResult{14, 6, 6}));
    a.trace().cigar_eq(q, r, res.query_idx, res.reference_idx, cigar);
    CHECK(cigar.to_string() == "3=2X1=");
    CHECK(cigar.len() == 3 && (cigar.get(1) == OpLen{Operation::X, 2}));
  }
  {
    auto r = PaddedBytes::from_bytes<AAMatrix>("AAAA", 16), q = PaddedBytes::from_bytes<AAMatrix>("AAA", 16);
    a.align(q, r, BLOSUM62(), g, {16, 16}, 0);
    const AlignResult res = a.res();
    CHECK((res == AlignResult{1, 3, 4}));
    a.trace().cigar(res.query_idx, res.reference_idx, cigar);
    CHECK(cigar.to_string() == "3M1D");
  }
  {
    auto r = PaddedBytes::from_bytes<NucMatrix>("TTAAAAAAATTTTTTTTTTTT", 16), q = PaddedBytes::from_bytes<NucMatrix>("TTTTTTTTAAAAAAATTTTTTTTT", 16);
    a.align(q, r, NW1(), g2, {16, 16}, 0);
    const AlignResult res = a.res();
    CHECK((res == AlignResult{7, 24, 21}));
    a.trace().cigar(res.query_idx, res.reference_idx, cigar);
    CHECK(cigar.to_string() == "2M6I16M3D");
  }
  Block<true, false> b(100, 100, 32);
  auto q = PaddedBytes::from_bytes<NucMatrix>("AAAAAAAAATTGCGCT", 32), r = PaddedBytes::from_bytes<NucMatrix>("AAAAAAAAAGCGC", 32);
  b.align(q, r, NW1(), g2, {32, 32}, 0);
  AlignResult res = b.res();
  CHECK((res == AlignResult{8, 16, 13}));
  b.trace().cigar_eq(q, r, res.query_idx, res.reference_idx, cigar);
  CHECK(cigar.to_string() == "9=2I4=1I");
  const NucMatrix m = NucMatrix::new_simple(2, -1);
  b.align(q, r, m, g3, {32, 32}, 0);
  res = b.res();
  CHECK((res == AlignResult{14, 16, 13}));
  b.trace().cigar_eq(q, r, res.query_idx, res.reference_idx, cigar);
  CHECK(cigar.to_string() == "9=2I4=1I");
}

static void test_bytes() {              // scan_block.rs:2105-2120
  Block<false, false> a(100, 100, 16);
  CHECK(run<ByteMatrix>(a, "AAAAAA", "AAAaaA", BYTES1(), Gaps{-2, -1}, 16, 16, 0, 16).score == 2);
  CHECK(run<ByteMatrix>(a, "abdefg", "abcdefg", BYTES1(), Gaps{-2, -1}, 16, 16, 0, 16).score == 4);
}

static void test_profile() {            // scan_block.rs:2122-2168
  Block<false, false> a(100, 100, 16);
  auto q4 = PaddedBytes::from_bytes<AAMatrix>("AAAA", 16);
  { auto r = AAProfile::from_bytes("AAAA", 16, 1, -1, -1, 0, -1, -1); a.align_profile(q4, r, {16, 16}, 0); CHECK(a.res().score == 4); }
  { auto r = AAProfile::from_bytes("AATTAA", 16, 1, -1, -1, 0, -1, -1); a.align_profile(q4, r, {16, 16}, 0); CHECK(a.res().score == 1); }
  { auto r = AAProfile::from_bytes("AATTAA", 16, 1, -1, -1, -1, -1, -1); a.align_profile(q4, r, {16, 16}, 0); CHECK(a.res().score == 0); }
  Block<true, false> t(100, 100, 16);
  Cigar cigar(100, 100);
  auto q = PaddedBytes::from_bytes<AAMatrix>("TTTTTTTTAAAAAAATTTTTTTTT", 16);
  struct { int8_t open_C, close_C; int score; const char* cig; bool tweak; } cases[] = {
      {-1, 0, 7, "2M6I16M3D", false}, {-1, -1, 6, "2M6I16M3D", false}, {-2, -1, 6, "2M6I14M3D2M", true}};
  for (auto& c : cases) {
    auto r = AAProfile::from_bytes("TTAAAAAAATTTTTTTTTTTT", 16, 1, -1, c.open_C, c.close_C, -1, -1);
    if (c.tweak) { r.set_gap_close_C(17, -1); r.set_gap_close_C(19, 0); }
    t.align_profile(q, r, {16, 16}, 0);
    const AlignResult res = t.res();
    CHECK((res == AlignResult{c.score, 24, 21}));
    t.trace().cigar(res.query_idx, res.reference_idx, cigar);
    CHECK(cigar.to_string() == c.cig);
  }
}

template <class B>
static void local_case(B& blk, const char* q, const char* r, int x, AlignResult want, const char* cig) {
  Cigar cigar(100, 100);
  auto rp = PaddedBytes::from_bytes<NucMatrix>(r, 32), qp = PaddedBytes::from_bytes<NucMatrix>(q, 32);
  blk.align(qp, rp, NW1(), Gaps{-2, -1}, {32, 32}, x);
  const AlignResult res = blk.res();
  CHECK((res == want));
  blk.trace().cigar_eq(qp, rp, res.query_idx, res.reference_idx, cigar);
  CHECK(cigar.to_string() == cig);
}
static void test_local_and_free_query_gaps() {   // scan_block.rs:2170-2230
  Block<true, false, true, false> local(100, 100, 32);
  local_case(local, "CCCCCCCCCCAAAAAA", "TTTTAAAAAA", 0, {6, 16, 10}, "6=");
  Block<true, true, true, false> local_x(100, 100, 32);
  local_case(local_x, "CCCCCCCCCCAAAAAACCCCCCCCCCCC", "TTTTAAAAAATTTTTTT", 100, {6, 16, 10}, "6=");
  Block<true, false, false, true> q_start(100, 100, 32);
  local_case(q_start, "AAAAAA", "CCCCCCCCCCAAAAAA", 0, {6, 6, 16}, "6=");
  local_case(q_start, "AAAAAA", "CCCCCCCCCCAAATAA", 0, {4, 6, 16}, "3=1X2=");
  Block<true, false, false, false, true> q_end(100, 100, 32);
  local_case(q_end, "AAAAAA", "AAAAAACCCCCCCCCC", 0, {6, 6, 6}, "6=");
  local_case(q_end, "AAAAAA", "AAATAACCCCCCCCCC", 0, {4, 6, 6}, "3=1X2=");
}

static void test_doc_example_and_helpers() {     // lib.rs:8-35, lib.rs:109-111, scan_block.rs:884-902
  const size_t min_block = 32, max_block = 256;
  auto r = PaddedBytes::from_bytes<NucMatrix>("TTAAAAAAATTTTTTTTTTTT", max_block);
  auto q = PaddedBytes::from_bytes<NucMatrix>("TTTTTTTTAAAAAAATTTTTTTTT", max_block);
  Block<true, false> a(q.len(), r.len(), max_block);
  a.align(q, r, NW1(), Gaps{-2, -1}, {min_block, max_block}, 0);
  const AlignResult res = a.res();
  CHECK((res == AlignResult{7, 24, 21}));
  Cigar cigar(res.query_idx, res.reference_idx);
  a.trace().cigar_eq(q, r, res.query_idx, res.reference_idx, cigar);
  CHECK(cigar.to_string() == "2=6I16=3D");
  CHECK(percent_len(1000, 0.1f) == 128 && percent_len(10, 0.5f) == 32);
  CHECK(q.get(0) == 'Z' && q.get(1) == 'T' && q.len() == 24);
  auto used = a.align_exp(q, r, NW1(), Gaps{-2, -1}, {min_block, max_block}, 0, 7);
  CHECK(used.has_value() && *used == 32);
  CHECK(!a.align_exp(q, r, NW1(), Gaps{-2, -1}, {min_block, max_block}, 0, 1000).has_value());
  // preconditions keep the reference's messages (scan_block.rs:799, 1483)
  bool threw = false;
  try { Block<false, false> bad(10, 10, 24); } catch (const Error& e) { threw = std::string(e.what()) == "Block size must be a power of two!"; }
  CHECK(threw);
  // batch extension
  std::vector<std::string> cigs;
  auto out = align_batch<true, false>({"TTTTTTTTAAAAAAATTTTTTTTT", "AAAAAAAAATTGCGCT"}, {"TTAAAAAAATTTTTTTTTTTT", "AAAAAAAAAGCGC"}, NW1(),
                                      Gaps{-2, -1}, {32, 32}, 0, &cigs);
  CHECK((out[0] == AlignResult{7, 24, 21}) && (out[1] == AlignResult{8, 16, 13}));
  CHECK(cigs[0] == "2=6I16=3D" && cigs[1] == "9=2I4=1I");
}

static void test_containers() {          // scan_block.rs:1807-1822, scores.rs:48-106, 150-183, 532-538
  auto fwd = PaddedBytes::from_bytes<NucMatrix>("TTGCA", 16);
  auto rev = PaddedBytes::new_<NucMatrix>(5, 16);
  rev.set_bytes_rev<NucMatrix>("ACGTT", 16);
  bool same = fwd.len() == rev.len();
  for (size_t i = 0; i <= fwd.len() + 16 && same; i++) same = fwd.get(i) == rev.get(i);
  CHECK(same);
  CHECK(fwd.get(0) == 'Z' && fwd.get(6) == 'Z');
  auto aa = PaddedBytes::from_bytes<AAMatrix>("arn", 16);          // lower case is upper-cased, stored as c - 'A'
  CHECK(aa.get(0) == 26 && aa.get(1) == 0 && aa.get(2) == 17 && aa.get(3) == 13);
  AAMatrix m = AAMatrix::new_simple(3, -2);
  CHECK(m.get('A', 'A') == 3 && m.get('A', 'R') == -2);
  m.set('A', 'R', 7);
  CHECK(m.get('R', 'A') == 7 && m.get('a', 'r') == 7);
  CHECK(BLOSUM62().get('A', 'A') == 4 && BLOSUM62().get('W', 'W') == 11 && BLOSUM62().get('A', 'R') == -1);
  NucMatrix n = NucMatrix::new_simple(2, -3);
  CHECK(n.get('A', 'A') == 2 && n.get('A', 'C') == -3 && NW1().get('G', 'G') == 1 && NW1().get('G', 'T') == -1);
  AAProfile p(3, 16, -1);
  p.set_all("AR", {1, 2, 3, 4, 5, 6});
  CHECK(p.len() == 3 && p.get(1, 'A') == 1 && p.get(1, 'R') == 2 && p.get(3, 'R') == 6 && p.get(2, 'N') == -128);
  p.set_all_rev("AR", {1, 2, 3, 4, 5, 6});
  CHECK(p.get(3, 'A') == 1 && p.get(1, 'R') == 6 && p.get_gap_extend() == -1);
  // a profile built from the same numbers as a sequence aligns like the sequence (scan_block.rs:2122-2127)
  Block<false, false> a(100, 100, 16);
  auto q = PaddedBytes::from_bytes<AAMatrix>("AAAA", 16);
  auto prof = AAProfile::from_bytes("AAAA", 16, 1, -1, -1, 0, -1, -1);
  a.align_profile(q, prof, {16, 16}, 0);
  CHECK(a.res().score == 4);
  CHECK(a.align_profile_exp(q, prof, {16, 16}, 0, 4).has_value());
}

int main() {
  test_containers();
  test_no_x_drop();
  test_x_drop();
  test_trace();
  test_bytes();
  test_profile();
  test_local_and_free_query_gaps();
  test_doc_example_and_helpers();
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
