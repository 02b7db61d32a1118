// ba_gen.cpp -- deterministic synthetic (query, reference) pair generator (host only, no CUDA).
//
// The reference's benches draw their inputs from the external `simulate-seqs` crate
// (benches/rand_scan.rs:96-103, examples/nanopore_bench.rs:43-49), which is not available here, so
// the workloads of BASELINE.json are re-defined on this generator: a
// splitmix64-seeded xoshiro256** per pair, keyed by (seed, stream, pair id), so any shard of any
// config can be regenerated bit-identically on any rank, for the GPU and for the CPU oracle alike.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>

namespace {

struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  Rng(uint64_t seed, uint64_t stream, uint64_t id) {
    uint64_t x = seed * 0x2545f4914f6cdd1dull + stream * 0x9e3779b97f4a7c15ull + id * 0xd1342543de82ef95ull + 0x1234567;
    for (int i = 0; i < 4; i++) s[i] = splitmix(x);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint64_t below(uint64_t n) { return n ? (uint64_t)(uni() * (double)n) % n : 0; }
  double normal() { double u1 = uni(), u2 = uni(); if (u1 < 1e-300) u1 = 1e-300; return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2); }
  uint32_t poisson(double mean) { double L = exp(-mean), p = 1.0; uint32_t k = 0; do { k++; p *= uni(); } while (p > L); return k - 1; }
  uint32_t geometric(double mean) { if (mean <= 1.0) return 1; double p = 1.0 / mean; double u = uni(); if (u < 1e-300) u = 1e-300; return 1 + (uint32_t)(log(u) / log(1.0 - p)); }
};

}  // namespace

extern "C" {

// All knobs of one workload. Plain C struct (mirrored with ctypes in tests/workloads.py).
struct BaGenParams {
  int32_t alphabet;          // 0: DNA "ACGT", 1: protein (20 standard residues, Robinson-Robinson-like background)
  int32_t len_dist;          // 0: uniform [len_min, len_max], 1: log-normal(median = len_median, sigma = len_sigma) clipped
  uint32_t len_min, len_max;
  double len_median, len_sigma;
  int32_t k_edits;           // >= 0: exactly k single-base edits (sub/ins/del with prob 1/3 each) instead of rates
  double sub_rate, ins_rate, del_rate;       // per reference base (used when k_edits < 0)
  double id_min, id_max;     // if id_max > 0: per-pair substitution rate = 1 - U[id_min, id_max] (overrides sub_rate)
  double indel_max;          // if > 0: per-pair ins_rate = del_rate = U[0, indel_max] / 2
  double long_indel_mean;    // Poisson mean of long indels per pair
  double long_indel_len;     // geometric mean length of a long indel
  double big_indel_prob;     // probability of one extra big indel
  uint32_t big_indel_min, big_indel_max;
  uint32_t suffix_len;       // unrelated random suffix appended to both (makes X-drop stop; nanopore_bench.rs:45-48)
};

static const char kDna[4] = {'A', 'C', 'G', 'T'};
static const char kAa[20] = {'A', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'K', 'L', 'M', 'N', 'P', 'Q', 'R', 'S', 'T', 'V', 'W', 'Y'};
// background frequencies x1000 (sum 1000), order of kAa
static const int kAaFreq[20] = {78, 19, 54, 63, 39, 74, 22, 52, 57, 90, 22, 45, 52, 43, 51, 71, 59, 64, 13, 32};

static inline uint8_t draw(Rng& g, int alphabet) {
  if (alphabet == 0) return (uint8_t)kDna[g.next() >> 62];
  int t = (int)g.below(1000), acc = 0;
  for (int i = 0; i < 20; i++) { acc += kAaFreq[i]; if (t < acc) return (uint8_t)kAa[i]; }
  return (uint8_t)'A';
}

static void gen_pair(const BaGenParams& P, uint64_t seed, uint64_t stream, uint64_t id, std::vector<uint8_t>& q, std::vector<uint8_t>& r) {
  Rng g(seed, stream, id);
  uint32_t len;
  if (P.len_dist == 1) {
    double v = P.len_median * exp(P.len_sigma * g.normal());
    if (v < P.len_min) v = P.len_min;
    if (v > P.len_max) v = P.len_max;
    len = (uint32_t)v;
  } else {
    len = P.len_min + (uint32_t)g.below((uint64_t)P.len_max - P.len_min + 1);
  }
  r.resize(len);
  for (uint32_t i = 0; i < len; i++) r[i] = draw(g, P.alphabet);
  q.clear();
  q.reserve(len + len / 4 + 64);
  if (P.k_edits >= 0) {
    q.assign(r.begin(), r.end());
    for (int e = 0; e < P.k_edits; e++) {
      const uint64_t kind = g.below(3);
      if (kind == 0 && !q.empty()) {
        q[g.below(q.size())] = draw(g, P.alphabet);
      } else if (kind == 1) {
        q.insert(q.begin() + (ptrdiff_t)g.below(q.size() + 1), draw(g, P.alphabet));
      } else if (!q.empty()) {
        q.erase(q.begin() + (ptrdiff_t)g.below(q.size()));
      }
    }
  } else {
    double sub = P.sub_rate, ins = P.ins_rate, del = P.del_rate;
    if (P.id_max > 0) sub = 1.0 - (P.id_min + (P.id_max - P.id_min) * g.uni());
    if (P.indel_max > 0) { const double t = P.indel_max * g.uni(); ins = del = t / 2; }
    // positions of long indels
    std::vector<std::pair<uint32_t, int32_t>> ev;  // (position, +len insertion / -len deletion)
    const uint32_t nl = P.long_indel_mean > 0 ? g.poisson(P.long_indel_mean) : 0;
    for (uint32_t e = 0; e < nl; e++) {
      const int32_t L = (int32_t)g.geometric(P.long_indel_len);
      ev.push_back({(uint32_t)g.below(len + 1), (g.next() & 1) ? L : -L});
    }
    if (P.big_indel_prob > 0 && g.uni() < P.big_indel_prob) {
      const int32_t L = (int32_t)(P.big_indel_min + g.below((uint64_t)P.big_indel_max - P.big_indel_min + 1));
      ev.push_back({(uint32_t)g.below(len + 1), (g.next() & 1) ? L : -L});
    }
    std::sort(ev.begin(), ev.end());
    size_t e = 0;
    uint32_t skip = 0;
    for (uint32_t i = 0; i <= len; i++) {
      while (e < ev.size() && ev[e].first == i) {
        if (ev[e].second > 0) for (int32_t t = 0; t < ev[e].second; t++) q.push_back(draw(g, P.alphabet));
        else skip += (uint32_t)(-ev[e].second);
        e++;
      }
      if (i == len) break;
      if (skip) { skip--; continue; }
      const double u = g.uni();
      if (u < del) continue;
      if (u < del + ins) { q.push_back(draw(g, P.alphabet)); q.push_back(r[i]); continue; }
      if (u < del + ins + sub) { uint8_t c = draw(g, P.alphabet); q.push_back(c); continue; }
      q.push_back(r[i]);
    }
  }
  for (uint32_t t = 0; t < P.suffix_len; t++) r.push_back(draw(g, P.alphabet));
  for (uint32_t t = 0; t < P.suffix_len; t++) q.push_back(draw(g, P.alphabet));
}

// pass 1: lengths of pairs [first, first+n)
void ba_gen_lengths(const BaGenParams* P, uint64_t seed, uint64_t stream, uint64_t first, uint64_t n,
                    uint32_t* q_len, uint32_t* r_len, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  auto work = [&](int t) {
    std::vector<uint8_t> q, r;
    for (uint64_t k = n * t / n_threads; k < n * (t + 1) / n_threads; k++) {
      gen_pair(*P, seed, stream, first + k, q, r);
      q_len[k] = (uint32_t)q.size(); r_len[k] = (uint32_t)r.size();
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}
// pass 2: bytes, written at the given offsets (n+1 entries each, from the lengths of pass 1)
void ba_gen_fill(const BaGenParams* P, uint64_t seed, uint64_t stream, uint64_t first, uint64_t n,
                 uint8_t* q_arena, const uint64_t* q_off, uint8_t* r_arena, const uint64_t* r_off, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  auto work = [&](int t) {
    std::vector<uint8_t> q, r;
    for (uint64_t k = n * t / n_threads; k < n * (t + 1) / n_threads; k++) {
      gen_pair(*P, seed, stream, first + k, q, r);
      if (!q.empty()) memcpy(q_arena + q_off[k], q.data(), q.size());
      if (!r.empty()) memcpy(r_arena + r_off[k], r.data(), r.size());
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}

}  // extern "C"
