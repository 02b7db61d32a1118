// ba_scalar.cpp -- TEST INFRASTRUCTURE: a second, independently written CPU restatement of the reference's adaptive
// block aligner, used only to cross-check oracle/ba_oracle.cpp (tests/test_oracle_scalar.py).
//
// ba_oracle.cpp is the literal restatement: the reference's own _mm256_* intrinsics, vector by vector. This file is
// the opposite on purpose: plain scalar C++, one DP cell at a time, no intrinsics, no vectors. The 16-lane prefix scan
// of the reference (src/avx2.rs:297-338) appears here only through its closed form
//     T[r] = max(T[r-1] + ext, x[r]),  T[top-1] = 0,     R[r] = max(T[r], ph[r mod 16]),
//     ph[k] = (k mod 8 + 1) * ext for k not in {7, 15},  ph[7] = 12 * ext,  ph[15] = none
// (the zeros shifted in by _mm256_slli_si256 act as extra candidates), and the per-lane
// argmax of the X-drop mode (src/scan_block.rs:1194-1201, src/avx2.rs:271-274) through its definition: for each AVX
// lane class (row mod 16) the last cell, in column-major order, that equals the running maximum of the class; the
// lowest class holding the block maximum wins. Everything else follows the step loop of align_core
// (src/scan_block.rs:94-595) as summarised in SURVEY.md Appendix A (A1-A9). No trace / CIGAR here: the walk-back is
// covered by the CIGAR re-scoring properties (tests/test_properties.py). FREE_QUERY_END_GAPS is not restated.
//
// Every reference unit test has min == max block size, so grow / shrink / checkpoint restore are pinned by nothing
// upstream; two restatements written against the same source in two different styles that agree on results AND on
// every step (direction, position, block size, offset, block maximum, border maxima) are the hedge.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace {

constexpr int L = 16, ZERO = 1 << 14, STEP = 8;
enum { F_TRACE = 1, F_XDROP = 2, F_LOCAL = 4, F_FQS = 8, F_FQE = 16 };
enum { K_NUC = 0, K_AA = 1, K_BYTE = 2, K_PROFILE = 3 };
enum { RIGHT = 0, DOWN = 1, GROW = 2 };

inline int sat(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

struct Step { int32_t dir; uint32_t i, j, block_size; int32_t off; int16_t max, right_max, down_max; };
struct Result { int32_t score; size_t query_idx, reference_idx; };

struct Scoring {
  int kind;
  const int8_t* matrix;                     // K_NUC: 8 x 16, K_AA: 27 x 32, K_BYTE: {match, mismatch}
  int open, ext;
  // profile (src/scores.rs:454-468): row per position
  const int8_t* pos_aa; const int16_t *open_C, *close_C, *open_R;
  const uint8_t *q, *r;                     // padded: index 0 is the pad byte
};

// substitution score of DP cell (row i, column j); profiles: the profile is the reference (columns)
inline int score_of(const Scoring& s, size_t i, size_t j) {
  const int a = s.q[i];
  if (s.kind == K_PROFILE) return s.pos_aa[j * 32 + a];
  const int b = s.r[j];
  if (s.kind == K_NUC) return s.matrix[(b & 7) * 16 + (a & 15)];        // src/scores.rs:195-209: row = reference byte
  if (s.kind == K_AA) return s.matrix[b * 32 + (a & 31)];               // src/scores.rs:110-127
  return a == b ? s.matrix[0] : s.matrix[1];                            // src/scores.rs:263-267
}

struct BlockMax { int max; int arg_v, arg_c; };   // arg: vector-direction index and column index of the argmax

// One rectangle. `right`: vectors run along rows (query), columns are reference positions; otherwise transposed.
// AD / AC: the border along the vector direction (in: previous column, already re-offset; out: last column);
// OD / OR: bottom "row" out, one entry per computed column. Returns the block maximum and its argmax under the
// reference's lane order. `ncols` <= W: columns computed (global-mode early break, src/scan_block.rs:1216-1224).
BlockMax place(const Scoring& sc, bool right, size_t vec_base, size_t col_base, int H, int W, int ncols, int16_t* AD, int16_t* AC,
               int16_t* OD, int16_t* OR_, int corner, int rz, bool local, bool fqs) {
  int lane_max[L], lane_v[L], lane_c[L];
  for (int k = 0; k < L; k++) { lane_max[k] = 0; lane_v[k] = 0; lane_c[k] = 0; }    // D_max starts at MIN = 0
  std::vector<int> Dp(H), Cp(H), Dn(H);
  for (int r = 0; r < H; r++) { Dp[r] = AD[r]; Cp[r] = AC[r]; }
  const bool prof = sc.kind == K_PROFILE;
  for (int c = 0; c < ncols; c++) {
    int T = 0;                                        // R01 = MIN: nothing enters the rectangle from above
    for (int r = 0; r < H; r++) {
      const size_t vi = vec_base + r, ci = col_base + c;
      const size_t row = right ? vi : ci, col = right ? ci : vi;          // DP-matrix coordinates of this cell
      int s = score_of(sc, row, col);
      int d00 = r == 0 ? (c == 0 ? corner : 0) : Dp[r - 1];              // only the first column sees the corner
      // cell (0, 0) is relative zero; FREE_QUERY_START_GAPS: all of row 0 (src/scan_block.rs:1130-1132)
      if (r == 0 && ((vi == 0 && ci == 0 && !local) || (fqs && right && vi == 0))) { d00 = rz; s = 0; }
      int open_c, close_c = 0, open_r, close_r = 0, ext = sc.ext;
      if (!prof) { open_c = sc.open; open_r = sc.open - sc.ext; }
      else if (right) { open_c = sc.open_C[col] + ext; close_c = sc.close_C[col]; open_r = sc.open_R[col]; }
      else { open_c = sc.open_R[col] + ext; open_r = sc.open_C[col]; close_r = sc.close_C[col]; }   // src/scan_block.rs:671-676
      const int c11 = std::max(sat(Cp[r] + ext), sat(Dp[r] + open_c));
      const int c11_end = prof && right ? sat(c11 + close_c) : c11;
      int d = std::max(sat(d00 + s), c11_end);
      if (local) d = std::max(d, rz);
      const int x = sat(d + open_r);
      T = std::max(T + ext, x);                                           // >= x >= -32768: no clamp needed
      const int m = r % L;
      int R = T;
      if (m != 15) R = std::max(R, (m == 7 ? 12 : (m % 8 + 1)) * ext);    // the scan's phantom candidates
      if (prof && !right) R = sat(R + close_r);
      d = std::max(d, R);
      Dn[r] = d; Cp[r] = c11;
      if (d >= lane_max[m]) { lane_max[m] = d; lane_v[m] = r - m; lane_c[m] = c; }   // last cell equal to the running max
      if (r == H - 1) { OD[c] = (int16_t)d; OR_[c] = (int16_t)T; }
    }
    Dp.swap(Dn);
  }
  if (ncols > 0) for (int r = 0; r < H; r++) { AD[r] = (int16_t)Dp[r]; AC[r] = (int16_t)Cp[r]; }
  BlockMax bm{0, 0, 0};
  for (int k = 0; k < L; k++) bm.max = std::max(bm.max, lane_max[k]);
  for (int k = 0; k < L; k++) if (lane_max[k] == bm.max) { bm.arg_v = lane_v[k] + k; bm.arg_c = lane_c[k]; break; }   // lowest lane
  return bm;
}

struct Aligner {
  Scoring sc; size_t qlen, rlen; int min_size, max_size, x_drop, flags;
  std::vector<int16_t> Dc, Cc, Dr, Rr, kDc, kCc, kDr, kRr, t1, t2;
  Result res; uint64_t cells = 0; std::vector<Step>* log = nullptr;

  static void offset(std::vector<int16_t>& a, int n, int add) { for (int k = 0; k < n; k++) a[k] = (int16_t)sat(a[k] + add); }
  // slide by STEP, re-offset the kept part, append the fresh values un-offset; returns old b1[STEP-1] + add (src/scan_block.rs:1040-1061)
  static int shift(std::vector<int16_t>& b1, std::vector<int16_t>& b2, const std::vector<int16_t>& f1, const std::vector<int16_t>& f2, int B, int add) {
    const int corner = sat(b1[STEP - 1] + add);
    for (int k = 0; k + STEP < B; k++) { b1[k] = (int16_t)sat(b1[k + STEP] + add); b2[k] = (int16_t)sat(b2[k + STEP] + add); }
    for (int k = 0; k < STEP; k++) { b1[B - STEP + k] = f1[k]; b2[B - STEP + k] = f2[k]; }
    return corner;
  }
  static int max_n(const std::vector<int16_t>& a, int from, int n) { int m = -32768; for (int k = 0; k < n; k++) m = std::max<int>(m, a[from + k]); return m; }

  void run() {
    const bool xdrop = flags & F_XDROP, local = flags & F_LOCAL, fqs = flags & F_FQS;
    const int ms = max_size;
    for (auto* v : {&Dc, &Cc, &Dr, &Rr, &kDc, &kCc, &kDr, &kRr}) v->assign(ms, 0);      // Allocated::clear: MIN = 0
    t1.assign(ms, 0); t2.assign(ms, 0);
    size_t i = 0, j = 0, i_ck = 0, j_ck = 0, best_i = 0, best_j = 0;
    int B = min_size, prev_size = 0, dir = GROW, prev_dir = GROW;
    int off = 0, off_max = 0, best_max = 0, off_ck = 0, y_iter = 0, x_iter = 0, D_corner = 0;
    for (;;) {
      const int prev_off = off;
      BlockMax bm{0, 0, 0}, gm{0, 0, 0};
      int off_add = 0;
      if (dir == GROW) D_corner = 0; else { off = off_max; off_add = sat(prev_off - off); }
      const int rz = sat(ZERO - off);
      auto early = [&](bool right, size_t vec_base, size_t col_base, int H, int W) {
        int n = W;
        const size_t vec_len = right ? qlen : rlen, col_len = right ? rlen : qlen;
        if (!xdrop && vec_base + H > vec_len) {
          long lim = (long)col_len - (long)col_base; if (lim < 0) lim = 0;
          if (lim + 1 < n) n = (int)lim + 1;
        }
        return n;
      };
      if (dir == RIGHT) {
        offset(Dc, B, off_add); offset(Cc, B, off_add);
        const int corner = prev_dir == DOWN ? sat(D_corner + off_add) : 0;
        const size_t cb = j + B - STEP;
        bm = place(sc, true, i, cb, B, STEP, early(true, i, cb, B, STEP), Dc.data(), Cc.data(), t1.data(), t2.data(), corner, rz, local, fqs);
        cells += (uint64_t)B * STEP;
        D_corner = shift(Dr, Rr, t1, t2, B, off_add);
      } else if (dir == DOWN) {
        offset(Dr, B, off_add); offset(Rr, B, off_add);
        const int corner = prev_dir == RIGHT ? sat(D_corner + off_add) : 0;
        const size_t cb = i + B - STEP;
        bm = place(sc, false, j, cb, B, STEP, early(false, j, cb, B, STEP), Dr.data(), Rr.data(), t1.data(), t2.data(), corner, rz, local, fqs);
        cells += (uint64_t)B * STEP;
        D_corner = shift(Dc, Cc, t1, t2, B, off_add);
      } else {
        // grow (src/scan_block.rs:247-329): down part, then right part; the checkpoint borders are overwritten afterwards
        const int W = B - prev_size;
        if (prev_size > 0) {
          gm = place(sc, false, j, i + prev_size, prev_size, W, early(false, j, i + prev_size, prev_size, W), Dr.data(), Rr.data(),
                     Dc.data() + prev_size, Cc.data() + prev_size, 0, rz, local, fqs);
          cells += (uint64_t)prev_size * W;
        }
        bm = place(sc, true, i, j + prev_size, B, W, early(true, i, j + prev_size, B, W), Dc.data(), Cc.data(), Dr.data() + prev_size,
                   Rr.data() + prev_size, 0, rz, local, fqs);
        cells += (uint64_t)B * W;
        for (int k = 0; k < B; k++) { kDc[k] = Dc[k]; kCc[k] = Cc[k]; kDr[k] = Dr[k]; kRr[k] = Rr[k]; }   // B entries, like the reference
      }
      int right_max = max_n(Dc, 0, STEP), down_max = max_n(Dr, 0, STEP);
      const int this_dir = dir;
      prev_dir = dir;
      const int mx = std::max(bm.max, gm.max);
      off_max = off + mx - ZERO;
      y_iter++;
      bool grow_no_max = this_dir == GROW;
      if (log) log->push_back(Step{this_dir, (uint32_t)i, (uint32_t)j, (uint32_t)B, off, (int16_t)mx, (int16_t)right_max, (int16_t)down_max});
      if (off_max > best_max) {
        if (xdrop) {
          if (this_dir == RIGHT) { best_i = i + bm.arg_v; best_j = j + (B - STEP) + bm.arg_c; }
          else if (this_dir == DOWN) { best_i = i + (B - STEP) + bm.arg_c; best_j = j + bm.arg_v; }
          else if (bm.max >= gm.max) { best_i = i + bm.arg_v; best_j = j + prev_size + bm.arg_c; }
          else { best_i = i + prev_size + gm.arg_c; best_j = j + gm.arg_v; }
        }
        if (B < max_size) {
          i_ck = i; j_ck = j; off_ck = off; grow_no_max = false;
          for (int k = 0; k < B; k++) { kDc[k] = Dc[k]; kCc[k] = Cc[k]; kDr[k] = Dr[k]; kRr[k] = Rr[k]; }
        }
        best_max = off_max; y_iter = 0;
      }
      if (xdrop) {
        if (off_max < best_max - x_drop) { if (x_iter < 1) x_iter++; else break; }
        else x_iter = 0;
      }
      if (i + B > qlen && j + B > rlen) break;
      if (j + B > rlen) { i += STEP; dir = DOWN; continue; }
      if (i + B > qlen) { j += STEP; dir = RIGHT; continue; }
      if (B * 2 <= max_size && (y_iter > B / STEP - 1 || grow_no_max)) {
        prev_size = B; B *= 2; dir = GROW; i = i_ck; j = j_ck; off = off_ck;
        for (int k = 0; k < prev_size; k++) { Dc[k] = kDc[k]; Cc[k] = kCc[k]; Dr[k] = kDr[k]; Rr[k] = kRr[k]; }
        y_iter = 0;
        continue;
      }
      if (B > min_size && y_iter == 0 && std::max(max_n(Dr, B - 2, 2), max_n(Dc, B - 2, 2)) >= mx) {
        // shrink (src/scan_block.rs:505-549): keep the bottom / right halves of the borders
        prev_dir = GROW;
        const int nb = B / 2;
        for (int k = 0; k < nb; k++) { Dc[k] = Dc[k + nb]; Cc[k] = Cc[k + nb]; Dr[k] = Dr[k + nb]; Rr[k] = Rr[k + nb]; }
        B = nb; i += nb; j += nb;
        i_ck = i; j_ck = j; off_ck = off;
        for (int k = 0; k < nb; k++) { kDc[k] = Dc[k]; kCc[k] = Cc[k]; kDr[k] = Dr[k]; kRr[k] = Rr[k]; }
        right_max = max_n(Dc, 0, STEP); down_max = max_n(Dr, 0, STEP);
        y_iter = 0;
      }
      if (down_max > right_max) { i += STEP; dir = DOWN; } else { j += STEP; dir = RIGHT; }
    }
    if (xdrop) { res.score = best_max; res.query_idx = best_i; res.reference_idx = best_j; }
    else {
      const int v = (dir == RIGHT || dir == GROW) ? Dc[qlen - i] : Dr[rlen - j];
      res.score = off + v - ZERO; res.query_idx = qlen; res.reference_idx = rlen;
    }
  }
};

}  // namespace

extern "C" {

// kind: 0 nucleotide, 1 amino acid, 2 byte; q / r padded like PaddedBytes (index 0 = pad). Returns 0, or 1 for arguments
// this restatement does not cover (TRACE is ignored; FREE_QUERY_END_GAPS is refused).
int sca_align(const uint8_t* q, size_t qlen, const uint8_t* r, size_t rlen, int kind, const int8_t* matrix, int open, int ext,
              size_t min_size, size_t max_size, int x_drop, int flags, int32_t* score, size_t* query_idx, size_t* reference_idx,
              uint64_t* cells, Step* log, size_t log_cap, size_t* log_n) {
  if (flags & F_FQE) return 1;
  Aligner a;
  memset(&a.sc, 0, sizeof(a.sc));
  a.sc.kind = kind; a.sc.matrix = matrix; a.sc.open = open; a.sc.ext = ext; a.sc.q = q; a.sc.r = r;
  a.qlen = qlen; a.rlen = rlen; a.min_size = (int)std::max<size_t>(min_size, L); a.max_size = (int)std::max<size_t>(max_size, L);
  a.x_drop = x_drop; a.flags = flags;
  std::vector<Step> lg;
  if (log) a.log = &lg;
  a.run();
  *score = a.res.score; *query_idx = a.res.query_idx; *reference_idx = a.res.reference_idx;
  if (cells) *cells = a.cells;
  if (log) { const size_t n = std::min(lg.size(), log_cap); memcpy(log, lg.data(), n * sizeof(Step)); }
  if (log_n) *log_n = lg.size();
  return 0;
}

// sequence-to-profile: the profile's arrays as the reference lays them out (pos_aa [len + pad][32], three i16 gap arrays)
int sca_align_profile(const uint8_t* q, size_t qlen, const int8_t* pos_aa, const int16_t* open_C, const int16_t* close_C,
                      const int16_t* open_R, size_t plen, int ext, size_t min_size, size_t max_size, int x_drop, int flags,
                      int32_t* score, size_t* query_idx, size_t* reference_idx, uint64_t* cells, Step* log, size_t log_cap, size_t* log_n) {
  if (flags & F_FQE) return 1;
  Aligner a;
  memset(&a.sc, 0, sizeof(a.sc));
  a.sc.kind = K_PROFILE; a.sc.ext = ext; a.sc.q = q; a.sc.pos_aa = pos_aa; a.sc.open_C = open_C; a.sc.close_C = close_C; a.sc.open_R = open_R;
  a.qlen = qlen; a.rlen = plen; a.min_size = (int)std::max<size_t>(min_size, L); a.max_size = (int)std::max<size_t>(max_size, L);
  a.x_drop = x_drop; a.flags = flags;
  std::vector<Step> lg;
  if (log) a.log = &lg;
  a.run();
  *score = a.res.score; *query_idx = a.res.query_idx; *reference_idx = a.res.reference_idx;
  if (cells) *cells = a.cells;
  if (log) { const size_t n = std::min(lg.size(), log_cap); memcpy(log, lg.data(), n * sizeof(Step)); }
  if (log_n) *log_n = lg.size();
  return 0;
}

}  // extern "C"
