// ba_kernel.cuh -- the adaptive block aligner on one warp.
//
// One warp owns one alignment. Lane l owns R consecutive rows of the current chunk of the block
// (R = block/32 capped at 8; larger blocks are swept in 256-row chunks). A rectangle of the DP matrix
// is computed column by column; inside a column the vertical-gap table obeys
//     T[r] = max(T[r-1] + ext, x[r]),  x[r] = D[r] + (open - ext)
// which is a max-plus prefix scan: lane-local chain over the R rows, Kogge-Stone over the 32 lane
// aggregates with warp shuffles, carry applied back per row. All scores are 16-bit values held in
// 32-bit registers with the reference's saturating-i16 semantics reproduced exactly (see notes).
//
// What is reproduced, with reference file:line (all into /root/reference/src/):
//   align_core state machine ............ scan_block.rs:94-595
//   place_block / place_block_profile ... scan_block.rs:1083-1228 / 612-783   -> place_rect()
//   just_offset / shift_and_offset ...... scan_block.rs:1003-1061
//   prefix scan incl. "phantom" terms .... avx2.rs:297-338 (closed form: SURVEY.md section 8a row 3)
//   Trace + cigar_core ................... scan_block.rs:1344-1672 (own packed layout, same walk)
//   score lookups ........................ scores.rs:121-127, 204-209, 263-267, 596-612
//
// Saturation notes. The reference adds with _mm256_adds_epi16. Here every value is an int holding
// an i16. Lower saturation is folded into DPX max(a+b,c) with c = -32768 (or with a third operand
// that is itself >= -32768). Upper saturation can only trigger where a non-negative quantity is
// added: diagonal + substitution score, border + off_add, profile gap_close; those get a min().
// Inside the scan no clamp is needed: T[r] >= x[r] >= -32768, and unclamped sums below -32768
// can never win a max against x[r].
#pragma once
#include "ba_types.h"
#include "ba_warp.cuh"

namespace ba {

constexpr int kNegBig = -(1 << 28);
constexpr int kI16Min = -32768;
constexpr int kI16Max = 32767;

BA_DEV int clamp16(int x) { return wp::imin(wp::imax(x, kI16Min), kI16Max); }
BA_DEV int sat_add(int a, int b) { return wp::imin(wp::viaddmax(a, b, kI16Min), kI16Max); }
BA_DEV int sat_add_lo(int a, int b) { return wp::viaddmax(a, b, kI16Min); }

// Per-warp working set. Borders live in shared memory as i16, exactly the arrays of the
// reference's `Allocated` (scan_block.rs:1252-1276) so that stale entries evolve identically.
struct WarpMem {
  int16_t *Dc, *Cc, *Dr, *Rr;       // D_col, C_col, D_row, R_row        [max_size]
  int16_t *t1, *t2;                 // temp_buf1/2                       [16]
  int16_t *kDc, *kCc, *kDr, *kRr;   // checkpoint copies (smem or global) [max_size]
  int32_t *misc;                    // [0..1]: A_old at the row above a chunk (double buffered)
  uint8_t *ecarry;                  // trace bit carried across chunks    [max_size] (only blocks > 256)
  const int8_t *mat;                // staged matrix (smem)
};

// ---------------------------------------------------------------------------------------------
// Scorers. `vec` = sequence along the vector direction (rows of a right-rectangle, columns of a
// down-rectangle), `col` = the other one. For sequence-sequence alignment right/down only swap
// the two (scan_block.rs:212-228).
// ---------------------------------------------------------------------------------------------
template <int KIND> struct SeqScorer {
  static constexpr int kKind = KIND;
  static constexpr bool kProfile = false;
  const int8_t* mat; const uint8_t* vec; const uint8_t* col;
  int go, ge;
  int b_match, b_mismatch;
  BA_DEV int row_token(uint32_t vidx) const {
    int b = vec[vidx];
    if (KIND == kNuc) return b & 15;       // pshufb index (avx2.rs:354-356); sequence bytes are < 0x80
    if (KIND == kAA) return b & 31;        // two-table lookup selected by bit 4 (avx2.rs:343-350)
    return b;
  }
  BA_DEV int col_token(uint32_t cidx) const {
    int c = col[cidx];
    if (KIND == kNuc) return (c & 7) * 16;  // scores.rs:195-197
    if (KIND == kAA) return c * 32;         // scores.rs:110-113
    return c;
  }
  BA_DEV int score(int ctok, int rtok) const {
    if (KIND == kByte) return ctok == rtok ? b_match : b_mismatch;  // scores.rs:263-267
    return (int)mat[ctok + rtok];
  }
};

// ---------------------------------------------------------------------------------------------
// place_rect: DP over a rectangle of H (vector direction) x W (sequential columns).
// ---------------------------------------------------------------------------------------------
struct RectArgs {
  uint32_t vec_base, col_base;   // start_i / start_j of place_block (scan_block.rs:1087-1088)
  int W, H, ncols;               // ncols <= W: columns actually computed (global-mode early break)
  int16_t *AD, *AC;              // D_col / C_col of place_block: read+written, H entries
  int16_t *OD, *OR_;             // D_row / R_row of place_block: bottom row out, W entries
  int corner;                    // D_corner for (row 0, col 0)
  int off_add;                   // just_offset folded into the load of AD/AC
  int rz;                        // relative_zero
  uint32_t* tw;                  // trace words of this rectangle (TRACE only)
};

// profile context handed to the rectangle (scan_block.rs:612-783)
struct ProfArgs {
  const ProfileDev* p;
  const uint8_t* q;              // padded query
};

template <int SCORING, bool RIGHT, int R, bool TRACE, bool XDROP, bool MULTI>
BA_DEV void place_rect_r(const SeqScorer<(SCORING == kProfile ? kAA : SCORING)>& sc, const ProfArgs& pa,
                         const RectArgs& a, const WarpMem& w, int& bv, unsigned& bkey) {
  constexpr bool PROF = SCORING == kProfile;
  constexpr int CH = 32 * R;
  const int lane = wp::lane_id();
  const int ge = PROF ? pa.p->gap_extend : sc.ge;
  const int go = sc.go;
  const int open_r = go - ge;  // gap_open - gap_extend (scan_block.rs:1143), always <= -1

  int kg[R], ph[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    kg[k] = (k + 1) * ge;
    const int m = (lane * R + k) & 15;  // AVX lane of this row (chunks are multiples of 16 rows)
    // zeros shifted in by _mm256_slli_si256 act as extra candidates (avx2.rs:321-337)
    ph[k] = (m == 15) ? kI16Min : (m == 7 ? 12 * ge : ((m & 7) + 1) * ge);
  }
  int dec[5];
#pragma unroll
  for (int s = 0; s < 5; s++) dec[s] = (R << s) * ge;
  const int lane_rg = lane * R * ge;

  const int nchunks = (a.H + CH - 1) / CH;
  const int ngroups = a.W >> 3;
  const bool origin = (a.vec_base == 0 && a.col_base == 0);  // scan_block.rs:1130

  for (int ch = 0; ch < nchunks; ch++) {
    const int v0 = ch * CH + lane * R;
    const bool act = v0 < a.H;
    const int rows_here = wp::imin(CH, a.H - ch * CH);
    const int last_lane = rows_here / R - 1;

    int D10[R], C10[R], rtok[R];
    // profile-down per-row data
    int p_openC[R], p_openR[R], p_closeR[R];
    const int8_t* p_row[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
      D10[k] = sat_add((int)a.AD[v0 + k], a.off_add);
      C10[k] = sat_add((int)a.AC[v0 + k], a.off_add);
      if (!PROF) {
        rtok[k] = sc.row_token(a.vec_base + v0 + k);
      } else if (RIGHT) {
        rtok[k] = pa.q[a.vec_base + v0 + k];   // query byte selects the entry of the position's row
      } else {
        const uint32_t idx = a.vec_base + v0 + k;                 // profile position (scan_block.rs:672-675)
        p_openC[k] = (int)wp::ldg(pa.p->gap_open_R + idx) + ge;   // get_gap_open_down_R
        p_openR[k] = (int)wp::ldg(pa.p->gap_open_C + idx);        // get_gap_open_down_C
        p_closeR[k] = (int)wp::ldg(pa.p->gap_close_C + idx);      // get_gap_close_down_C
        p_row[k] = pa.p->pos_aa + (size_t)idx * 32;
        rtok[k] = 0;
      }
    }
    int diag_carry = a.corner;
    if (MULTI) {
      // value of the old border at the row just above this chunk = diagonal input of column 0
      if (ch > 0) diag_carry = w.misc[ch & 1];
      if (lane == 31) w.misc[(ch + 1) & 1] = D10[R - 1];
      wp::syncwarp();
    }

    int mx[R];
    unsigned mcol[R];
#pragma unroll
    for (int k = 0; k < R; k++) { mx[k] = 0; mcol[k] = 0; }
    unsigned twd[R];
#pragma unroll
    for (int k = 0; k < R; k++) twd[k] = 0;

    for (int cg = 0; cg < a.ncols; cg += 8) {
      int Ttop[8], Dtop[8], etop[8];
#pragma unroll
      for (int c = 0; c < 8; c++) { Ttop[c] = 0; Dtop[c] = 0; etop[c] = 0; }
      if (MULTI) {
        if (ch > 0) {
#pragma unroll
          for (int c = 0; c < 8; c++) {
            if (cg + c < a.ncols) {
              Ttop[c] = (int)a.OR_[cg + c];
              Dtop[c] = (int)a.OD[cg + c];
              if (TRACE) etop[c] = (int)w.ecarry[cg + c];
            }
          }
        }
        wp::syncwarp();
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (cg + c < a.ncols) {
          const int cidx = cg + c;
          // ---- per-column uniform data ----
          int ctok = 0, openC_col = 0, closeC_col = 0, openR_col = 0;
          const int8_t* prow_col = nullptr;
          if (!PROF) {
            ctok = sc.col_token(a.col_base + cidx);
          } else if (RIGHT) {
            const uint32_t idx = a.col_base + cidx;                   // scan_block.rs:658-661
            openC_col = (int)wp::ldg(pa.p->gap_open_C + idx) + ge;
            closeC_col = (int)wp::ldg(pa.p->gap_close_C + idx);
            openR_col = (int)wp::ldg(pa.p->gap_open_R + idx);
            prow_col = pa.p->pos_aa + (size_t)idx * 32;
          } else {
            ctok = pa.q[a.col_base + cidx];                            // query byte of this "column"
          }
          // ---- diagonal input of the lane's first row ----
          int up = wp::shfl_up(D10[R - 1], 1);
          if (lane == 0) up = diag_carry;
          diag_carry = Dtop[c];

          int dd[R], xx[R], c11[R], c11o[R], c11e[R], tt[R];
#pragma unroll
          for (int k = 0; k < R; k++) {
            int s;
            if (!PROF) s = sc.score(ctok, rtok[k]);
            else if (RIGHT) s = (int)wp::ldg(prow_col + rtok[k]);
            else s = (int)wp::ldg(p_row[k] + ctok);
            int d00 = (k == 0) ? up : D10[k - 1];
            if (origin && cidx == 0 && ch == 0 && k == 0 && lane == 0) { d00 = a.rz; s = 0; }
            const int oc = PROF ? (RIGHT ? openC_col : p_openC[k]) : go;
            c11o[k] = sat_add_lo(D10[k], oc);
            c11[k] = wp::viaddmax(C10[k], ge, c11o[k]);
            c11e[k] = c11[k];
            if (PROF && RIGHT) c11e[k] = sat_add(c11[k], closeC_col);   // C11_end (scan_block.rs:694)
            dd[k] = wp::imin(wp::viaddmax(d00, s, c11e[k]), kI16Max);
            const int orr = PROF ? (RIGHT ? openR_col : p_openR[k]) : open_r;
            xx[k] = sat_add_lo(dd[k], orr);
            tt[k] = (k == 0) ? xx[0] : wp::viaddmax(tt[k - 1], ge, xx[k]);
          }
          // ---- cross-lane max-plus scan of the lane aggregates ----
          int inc = tt[R - 1];
#pragma unroll
          for (int s = 0; s < 5; s++) {
            const int u = wp::shfl_up(inc, 1 << s);
            inc = wp::viaddmax(u, dec[s], inc);
          }
          int ex = wp::shfl_up(inc, 1);
          if (lane == 0) ex = kNegBig;
          const int cin = wp::viaddmax(Ttop[c], lane_rg, ex);

          int Dn[R], Tn[R];
          unsigned ebits = 0;
#pragma unroll
          for (int k = 0; k < R; k++) {
            Tn[k] = wp::viaddmax(cin, kg[k], tt[k]);
            const int Rv = wp::imax(Tn[k], ph[k]);
            int Rend = Rv;
            if (PROF && !RIGHT) Rend = sat_add(Rv, p_closeR[k]);
            Dn[k] = wp::imax(dd[k], Rend);
            if (TRACE) {
              unsigned nib = (Dn[k] == c11e[k] ? 1u : 0u) | (Dn[k] == Rend ? 2u : 0u);
              nib |= (c11[k] == c11o[k] ? 4u : 0u);
              ebits |= (Rv == xx[k] ? 1u : 0u) << k;
              twd[k] |= nib << (4 * c);
            }
            if (XDROP) {
              const int nm = wp::imax(mx[k], Dn[k]);
              if (Dn[k] == nm) mcol[k] = (unsigned)cidx + 1u;
              mx[k] = nm;
            } else {
              mx[0] = wp::imax(mx[0], act ? Dn[k] : 0);
            }
          }
          if (TRACE) {
            // "R gap at this row was opened at the row above" = e of the row above (scan_block.rs:1179-1181)
            const unsigned eb = wp::ballot(((ebits >> (R - 1)) & 1u) != 0);
            unsigned above = lane == 0 ? (unsigned)etop[c] : ((eb >> (lane - 1)) & 1u);
#pragma unroll
            for (int k = 0; k < R; k++) {
              const unsigned b3 = (k == 0) ? above : ((ebits >> (k - 1)) & 1u);
              twd[k] |= (b3 << 3) << (4 * c);
            }
            if (MULTI && lane == last_lane) w.ecarry[cidx] = (uint8_t)((ebits >> (R - 1)) & 1u);
          }
          // ---- bottom row of the rectangle (or of this chunk) ----
          if (lane == last_lane) {
            a.OD[cidx] = (int16_t)Dn[R - 1];
            a.OR_[cidx] = (int16_t)Tn[R - 1];
          }
#pragma unroll
          for (int k = 0; k < R; k++) { D10[k] = Dn[k]; C10[k] = c11[k]; }
        }
      }
      if (TRACE) {
        const int cgi = cg >> 3;
#pragma unroll
        for (int k = 0; k < R; k++) {
          a.tw[(((size_t)ch * ngroups + cgi) * R + k) * 32 + lane] = twd[k];
          twd[k] = 0;
        }
      }
      if (MULTI) wp::syncwarp();
    }

    if (act) {
#pragma unroll
      for (int k = 0; k < R; k++) {
        a.AD[v0 + k] = (int16_t)D10[k];
        a.AC[v0 + k] = (int16_t)C10[k];
      }
    }
    // merge the per-row trackers into the thread's best under the reference's tie-break order:
    // value desc, AVX lane (row mod 16) asc, column desc, row desc  (scan_block.rs:1194-1201, avx2.rs:271-274)
    if (XDROP) {
      if (act) {
#pragma unroll
        for (int k = 0; k < R; k++) {
          if (mcol[k] != 0) {
            const unsigned cls = (unsigned)((lane * R + k) & 15);
            const unsigned key = ((15u - cls) << 27) | (mcol[k] << 13) | (unsigned)(v0 + k);
            if (mx[k] > bv || (mx[k] == bv && key > bkey)) { bv = mx[k]; bkey = key; }
          }
        }
      }
    } else {
      bv = wp::imax(bv, mx[0]);
    }
  }
  wp::syncwarp();
}

template <int SCORING, bool RIGHT, bool TRACE, bool XDROP>
BA_DEV void place_rect(const SeqScorer<(SCORING == kProfile ? kAA : SCORING)>& sc, const ProfArgs& pa,
                       const RectArgs& a, const WarpMem& w, int& bv, unsigned& bkey) {
  bv = 0;                // D_max starts at MIN = 0 (scan_block.rs:1101)
  bkey = 15u << 27;      // "no cell": AVX lane 0 with argmax (0, 0)
  if (a.W == 0 || a.H == 0) return;   // scan_block.rs:1105-1107
  if (a.H <= 32) place_rect_r<SCORING, RIGHT, 1, TRACE, XDROP, false>(sc, pa, a, w, bv, bkey);
  else if (a.H == 64) place_rect_r<SCORING, RIGHT, 2, TRACE, XDROP, false>(sc, pa, a, w, bv, bkey);
  else if (a.H == 128) place_rect_r<SCORING, RIGHT, 4, TRACE, XDROP, false>(sc, pa, a, w, bv, bkey);
  else if (a.H == 256) place_rect_r<SCORING, RIGHT, 8, TRACE, XDROP, false>(sc, pa, a, w, bv, bkey);
  else place_rect_r<SCORING, RIGHT, 8, TRACE, XDROP, true>(sc, pa, a, w, bv, bkey);
}

// ---------------------------------------------------------------------------------------------
// Border helpers
// ---------------------------------------------------------------------------------------------
// shift_and_offset (scan_block.rs:1040-1061): slide by STEP, re-offset the kept part, append the
// 8 fresh values from temp (not re-offset). Returns old buf1[STEP-1] + off_add.
BA_DEV int shift_and_offset(int B, int16_t* b1, int16_t* b2, const int16_t* t1, const int16_t* t2, int off_add) {
  const int lane = wp::lane_id();
  const int corner = sat_add((int)b1[kStep - 1], off_add);
  wp::syncwarp();
  for (int base = 0; base < B; base += 32) {
    const int idx = base + lane;
    int v1 = 0, v2 = 0;
    if (idx < B) {
      if (idx < B - kStep) {
        v1 = sat_add((int)b1[idx + kStep], off_add);
        v2 = sat_add((int)b2[idx + kStep], off_add);
      } else {
        v1 = (int)t1[idx - (B - kStep)];
        v2 = (int)t2[idx - (B - kStep)];
      }
    }
    wp::syncwarp();
    if (idx < B) { b1[idx] = (int16_t)v1; b2[idx] = (int16_t)v2; }
    wp::syncwarp();
  }
  return corner;
}

BA_DEV void copy4(int n, int16_t* d0, int16_t* d1, int16_t* d2, int16_t* d3,
                  const int16_t* s0, const int16_t* s1, const int16_t* s2, const int16_t* s3, int src_off) {
  const int lane = wp::lane_id();
  wp::syncwarp();
  for (int base = 0; base < n; base += 32) {
    const int idx = base + lane;
    int16_t a = 0, b = 0, c = 0, d = 0;
    if (idx < n) { a = s0[idx + src_off]; b = s1[idx + src_off]; c = s2[idx + src_off]; d = s3[idx + src_off]; }
    wp::syncwarp();
    if (idx < n) { d0[idx] = a; d1[idx] = b; d2[idx] = c; d3[idx] = d; }
  }
  wp::syncwarp();
}

// prefix_max over the first STEP entries of two borders, suffix_max over the last two
// (scan_block.rs:1020-1032)
BA_DEV void border_maxes(const int16_t* Dc, const int16_t* Dr, int& right_max, int& down_max) {
  const int lane = wp::lane_id();
  wp::syncwarp();
  int v = kI16Min;
  if (lane < 8) v = (int)Dc[lane];
  else if (lane < 16) v = (int)Dr[lane - 8];
  right_max = wp::red_max(lane < 8 ? v : kI16Min);
  down_max = wp::red_max((lane >= 8 && lane < 16) ? v : kI16Min);
}
BA_DEV int shrink_max(const int16_t* Dc, const int16_t* Dr, int B) {
  const int lane = wp::lane_id();
  int v = kI16Min;
  if (lane < 2) v = (int)Dr[B - 2 + lane];
  else if (lane < 4) v = (int)Dc[B - 4 + lane];
  return wp::red_max(v);
}

// ---------------------------------------------------------------------------------------------
// Trace bookkeeping (stack of rectangles + packed words; scan_block.rs:1344-1463)
// ---------------------------------------------------------------------------------------------
struct TraceState {
  uint32_t* words; uint64_t words_cap;
  Rect* rects; uint32_t rects_cap;
  uint64_t widx; uint32_t ridx;
  uint64_t ck_widx; uint32_t ck_ridx;
  bool overflow;
};
BA_DEV uint64_t rect_words(int H, int W) {
  const int R = H <= 32 ? 1 : (H >= 256 ? 8 : H / 32);
  const int CH = 32 * R;
  const int nch = (H + CH - 1) / CH;
  return (uint64_t)nch * (uint64_t)(W >> 3) * R * 32;
}
BA_DEV uint32_t* trace_push(TraceState& ts, uint32_t row, uint32_t col, int W, int H, bool right) {
  const uint64_t need = rect_words(H, W);
  if (ts.ridx >= ts.rects_cap || ts.widx + need > ts.words_cap || (ts.widx + need) >> 32) { ts.overflow = true; return ts.words; }
  if (wp::lane_id() == 0) {
    Rect r; r.row = row; r.col = col; r.h = (uint16_t)H; r.w = (uint16_t)W; r.right = right ? 1u : 0u; r.word_off = (uint32_t)ts.widx;
    ts.rects[ts.ridx] = r;
  }
  uint32_t* p = ts.words + ts.widx;
  ts.widx += need;
  ts.ridx += 1;
  return p;
}

// ---------------------------------------------------------------------------------------------
// Traceback (Trace::cigar_core, scan_block.rs:1482-1672). Lane 0 walks from (i, j) back to the
// origin through the stack of rectangles; runs are produced reversed and run-length merged
// exactly like Cigar::add (cigar.rs:71-79), then written forward into the output stream.
// ---------------------------------------------------------------------------------------------
BA_DEV void traceback_walk(const uint32_t* words, const Rect* rects, uint32_t ridx, uint32_t i, uint32_t j,
                           const uint8_t* q, const uint8_t* r, bool eq, uint32_t* runs, uint32_t cap,
                           uint32_t& nruns, uint32_t& bad) {
  int table = 0;  // 0 = D, 1 = C, 2 = R
  uint32_t cur_op = 0, cur_len = 0;
  nruns = 0; bad = 0;
  while ((i > 0 || j > 0) && !bad) {
    Rect rc;
    for (;;) {   // find the newest rectangle containing (i, j) (scan_block.rs:1578-1590)
      if (ridx == 0) { bad = 1; break; }
      ridx--;
      rc = rects[ridx];
      if (i >= rc.row && j >= rc.col) break;
    }
    if (bad) break;
    const int H = rc.h, W = rc.w;
    const int R = H <= 32 ? 1 : (H >= 256 ? 8 : H / 32);
    const int CH = 32 * R;
    const int ngroups = W >> 3;
    const uint32_t* tw = words + rc.word_off;
    // The 64-entry OP_LUT (scan_block.rs:1508-1572) folded into branches. Bit 0 of trace/trace2
    // talks about the gap table that runs along the rectangle's sequential columns (C for a right
    // rectangle, R for a down rectangle), bit 1 about the table along its vectors.
    const int tabA = rc.right ? 1 : 2, tabB = rc.right ? 2 : 1;
    const uint32_t opA = rc.right ? 5u : 4u, opB = rc.right ? 4u : 5u;   // D : I
    while (i >= rc.row && j >= rc.col && (i > 0 || j > 0)) {
      const uint32_t v = rc.right ? i - rc.row : j - rc.col;
      const uint32_t c = rc.right ? j - rc.col : i - rc.row;
      const uint32_t ch = v / CH, ln = (v % CH) / R, k = v % R;
      const uint32_t word = tw[(((size_t)ch * ngroups + (c >> 3)) * R + k) * 32 + ln];
      const uint32_t nib = (word >> (4 * (c & 7))) & 15u;
      const uint32_t t = nib & 3u, t2 = nib >> 2;
      uint32_t op; int ntab;
      if (table == tabA) { op = opA; ntab = (t2 & 1u) ? 0 : tabA; }
      else if (table == tabB) { op = opB; ntab = (t2 & 2u) ? 0 : tabB; }
      else if (t == 0) { op = 1u; ntab = 0; }
      else if (t & 1u) { op = opA; ntab = (t2 & 1u) ? 0 : tabA; }
      else { op = opB; ntab = (t2 & 2u) ? 0 : tabB; }
      const uint32_t di = (op == 5u) ? 0u : 1u, dj = (op == 4u) ? 0u : 1u;
      if (op == 1u && eq) op = (q[i] == r[j]) ? 2u : 3u;   // scan_block.rs:1620-1628
      if ((di && i == 0) || (dj && j == 0)) { bad = 1; break; }
      i -= di; j -= dj; table = ntab;
      if (op == cur_op) cur_len++;
      else {
        if (cur_len) { if (nruns < cap) runs[nruns] = (cur_len << 4) | cur_op; nruns++; }
        cur_op = op; cur_len = 1;
      }
    }
  }
  if (cur_len) { if (nruns < cap) runs[nruns] = (cur_len << 4) | cur_op; nruns++; }
}

BA_DEV void emit_cigar(const Params& P, const uint32_t* words, const Rect* rects, uint32_t ridx, uint32_t qi, uint32_t rj,
                       const uint8_t* q, const uint8_t* r, bool eq, uint32_t* runs, DevResult& res) {
  const int lane = wp::lane_id();
  uint32_t nruns = 0, bad = 0;
  if (lane == 0) traceback_walk(words, rects, ridx, qi, rj, q, r, eq, runs, P.runs_per_warp, nruns, bad);
  nruns = (uint32_t)wp::shfl_idx((int)nruns, 0);
  bad = (uint32_t)wp::shfl_idx((int)bad, 0);
  res.cigar_n = 0; res.cigar_off = 0;
  if (bad || nruns > P.runs_per_warp) { res.status = (uint32_t)kCigarOverflow; return; }
  unsigned long long base = 0;
  if (lane == 0) base = wp::atomic_add64(P.cigar_used, (unsigned long long)nruns);
  const uint32_t blo = (uint32_t)wp::shfl_idx((int)(uint32_t)(base & 0xffffffffu), 0);
  const uint32_t bhi = (uint32_t)wp::shfl_idx((int)(uint32_t)(base >> 32), 0);
  base = ((unsigned long long)bhi << 32) | blo;
  if (base + nruns > P.cigar_cap) { res.status = (uint32_t)kCigarOverflow; return; }
  wp::syncwarp();
  for (uint32_t t = lane; t < nruns; t += 32) P.cigar_stream[base + t] = runs[nruns - 1 - t];
  res.cigar_n = nruns; res.cigar_off = base;
}

// ---------------------------------------------------------------------------------------------
// The per-alignment state machine (scan_block.rs:94-595)
// ---------------------------------------------------------------------------------------------
template <int SCORING, int FLAGS>
BA_DEV void align_pair(const Params& P, uint32_t pair, const WarpMem& w, TraceState& ts, uint32_t warp_global) {
  constexpr bool TRACE = (FLAGS & kTrace) != 0, XDROP = (FLAGS & kXDrop) != 0;
  constexpr bool PROF = SCORING == kProfile;
  const int lane = wp::lane_id();

  const uint8_t* q = P.seq + P.q_off[pair];
  const uint32_t qlen = P.q_len[pair];
  const uint8_t* r = nullptr;
  uint32_t rlen;
  ProfArgs pa; pa.p = nullptr; pa.q = q;
  if (PROF) { pa.p = P.profiles + pair; rlen = pa.p->len; }
  else { r = P.seq + P.r_off[pair]; rlen = P.r_len[pair]; }

  SeqScorer<(SCORING == kProfile ? kAA : SCORING)> sc;
  sc.mat = w.mat; sc.vec = q; sc.col = r; sc.go = P.gap_open; sc.ge = P.gap_extend;
  sc.b_match = sc.b_mismatch = 0;
  if (SCORING == kByte) { sc.b_match = (int)P.matrix[0]; sc.b_mismatch = (int)P.matrix[1]; }

  const int min_size = (int)P.min_size, max_size = (int)P.max_size;

  // Allocated::clear (scan_block.rs:1322-1339): every border starts at MIN = 0
  for (int idx = lane; idx < max_size; idx += 32) {
    w.Dc[idx] = 0; w.Cc[idx] = 0; w.Dr[idx] = 0; w.Rr[idx] = 0;
    w.kDc[idx] = 0; w.kCc[idx] = 0; w.kDr[idx] = 0; w.kRr[idx] = 0;
  }
  if (lane < 16) { w.t1[lane] = 0; w.t2[lane] = 0; }
  ts.widx = 0; ts.ridx = 0; ts.ck_widx = 0; ts.ck_ridx = 0; ts.overflow = false;
  wp::syncwarp();

  uint32_t si = 0, sj = 0;
  int best_max = 0; uint32_t best_i = 0, best_j = 0;
  int prev_dir = kGrow, dir = kGrow;
  int prev_size = 0, B = min_size;
  int off = 0, prev_off, off_max = 0;
  int y_drop_iter = 0, x_drop_iter = 0;
  uint32_t i_ckpt = 0, j_ckpt = 0; int off_ckpt = 0;
  int D_corner = 0;
  uint64_t cells = 0; uint32_t steps = 0;

  for (;;) {
    prev_off = off;
    int bv = 0, gbv = 0; unsigned bkey = 15u << 27, gbkey = 15u << 27;
    int right_max, down_max;
    steps++;
    RectArgs a;
    if (dir == kRight) {
      off = off_max;
      const int off_add = clamp16(prev_off - off);
      a.vec_base = si; a.col_base = sj + B - kStep; a.W = kStep; a.H = B;
      a.AD = w.Dc; a.AC = w.Cc; a.OD = w.t1; a.OR_ = w.t2;
      a.corner = (prev_dir == kDown) ? sat_add(D_corner, off_add) : 0;
      a.off_add = off_add; a.rz = clamp16(kZero - off);
      a.ncols = a.W;
      if (!XDROP && a.vec_base + a.H > qlen) { int lim = (int)rlen - (int)a.col_base; if (lim < 0) lim = 0; if (lim + 1 < a.ncols) a.ncols = lim + 1; }
      a.tw = nullptr;
      if (TRACE) a.tw = trace_push(ts, si, sj + B - kStep, kStep, B, true);
      cells += (uint64_t)kStep * B;
      sc.vec = q; sc.col = r;
      if (!ts.overflow) place_rect<SCORING, true, TRACE, XDROP>(sc, pa, a, w, bv, bkey);
      D_corner = shift_and_offset(B, w.Dr, w.Rr, w.t1, w.t2, off_add);
      border_maxes(w.Dc, w.Dr, right_max, down_max);
    } else if (dir == kDown) {
      off = off_max;
      const int off_add = clamp16(prev_off - off);
      a.vec_base = sj; a.col_base = si + B - kStep; a.W = kStep; a.H = B;
      a.AD = w.Dr; a.AC = w.Rr; a.OD = w.t1; a.OR_ = w.t2;
      a.corner = (prev_dir == kRight) ? sat_add(D_corner, off_add) : 0;
      a.off_add = off_add; a.rz = clamp16(kZero - off);
      a.ncols = a.W;
      if (!XDROP && a.vec_base + a.H > rlen) { int lim = (int)qlen - (int)a.col_base; if (lim < 0) lim = 0; if (lim + 1 < a.ncols) a.ncols = lim + 1; }
      a.tw = nullptr;
      if (TRACE) a.tw = trace_push(ts, si + B - kStep, sj, kStep, B, false);
      cells += (uint64_t)kStep * B;
      sc.vec = r; sc.col = q;
      if (!ts.overflow) place_rect<SCORING, false, TRACE, XDROP>(sc, pa, a, w, bv, bkey);
      D_corner = shift_and_offset(B, w.Dc, w.Cc, w.t1, w.t2, off_add);
      border_maxes(w.Dc, w.Dr, right_max, down_max);
    } else {
      D_corner = 0;
      const int grow_step = B - prev_size;
      // down part: rows si+prev.. x cols sj..sj+prev (scan_block.rs:262-278)
      a.vec_base = sj; a.col_base = si + prev_size; a.W = grow_step; a.H = prev_size;
      a.AD = w.Dr; a.AC = w.Rr; a.OD = w.Dc + prev_size; a.OR_ = w.Cc + prev_size;
      a.corner = 0; a.off_add = 0; a.rz = clamp16(kZero - off);
      a.ncols = a.W;
      if (!XDROP && a.vec_base + a.H > rlen) { int lim = (int)qlen - (int)a.col_base; if (lim < 0) lim = 0; if (lim + 1 < a.ncols) a.ncols = lim + 1; }
      a.tw = nullptr;
      if (TRACE) a.tw = trace_push(ts, si + prev_size, sj, grow_step, prev_size, false);
      cells += (uint64_t)grow_step * prev_size;
      sc.vec = r; sc.col = q;
      if (!ts.overflow) place_rect<SCORING, false, TRACE, XDROP>(sc, pa, a, w, gbv, gbkey);
      // right part: rows si..si+B x cols sj+prev..sj+B (scan_block.rs:289-305)
      a.vec_base = si; a.col_base = sj + prev_size; a.W = grow_step; a.H = B;
      a.AD = w.Dc; a.AC = w.Cc; a.OD = w.Dr + prev_size; a.OR_ = w.Rr + prev_size;
      a.ncols = a.W;
      if (!XDROP && a.vec_base + a.H > qlen) { int lim = (int)rlen - (int)a.col_base; if (lim < 0) lim = 0; if (lim + 1 < a.ncols) a.ncols = lim + 1; }
      if (TRACE) a.tw = trace_push(ts, si, sj + prev_size, grow_step, B, true);
      cells += (uint64_t)grow_step * B;
      sc.vec = q; sc.col = r;
      if (!ts.overflow) place_rect<SCORING, true, TRACE, XDROP>(sc, pa, a, w, bv, bkey);
      border_maxes(w.Dc, w.Dr, right_max, down_max);
      copy4(B, w.kDc, w.kCc, w.kDr, w.kRr, w.Dc, w.Cc, w.Dr, w.Rr, 0);   // scan_block.rs:315-322
      if (TRACE) { ts.ck_widx = ts.widx; ts.ck_ridx = ts.ridx; }
    }
    if (ts.overflow) break;

    const int this_dir = dir;
    prev_dir = dir;
    const int D_max_max = wp::red_max(bv);
    const int grow_max = wp::red_max(gbv);
    const int mx = wp::imax(D_max_max, grow_max);
    off_max = off + mx - kZero;
    y_drop_iter++;
    bool grow_no_max = (dir == kGrow);

    if (P.step_log && lane == 0) {
      const uint32_t n = *P.step_log_n;
      if (n < P.step_log_cap) {
        StepLog s; s.dir = dir; s.i = si; s.j = sj; s.block_size = (uint32_t)B; s.off = off;
        s.max = (int16_t)mx; s.right_max = (int16_t)right_max; s.down_max = (int16_t)down_max;
        P.step_log[n] = s;
      }
      *P.step_log_n = n + 1;
    }

    if (off_max > best_max) {
      if (XDROP) {
        // decode the argmax (scan_block.rs:370-404)
        const bool use_right = (this_dir != kGrow) || (D_max_max >= grow_max);
        const int m = use_right ? D_max_max : grow_max;
        const int tv = use_right ? bv : gbv;
        const unsigned tk = use_right ? bkey : gbkey;
        const unsigned key = wp::red_max_u(tv == m ? tk : 0u);
        const unsigned cp1 = (key >> 13) & 0x3fffu;
        uint32_t v = key & 0x1fffu, c = 0;
        if (cp1 == 0) v = 0; else c = cp1 - 1;
        if (this_dir == kRight) { best_i = si + v; best_j = sj + (uint32_t)(B - kStep) + c; }
        else if (this_dir == kDown) { best_i = si + (uint32_t)(B - kStep) + c; best_j = sj + v; }
        else if (use_right) { best_i = si + v; best_j = sj + (uint32_t)prev_size + c; }
        else { best_i = si + (uint32_t)prev_size + c; best_j = sj + v; }
      }
      if (B < max_size) {
        i_ckpt = si; j_ckpt = sj; off_ckpt = off;
        copy4(B, w.kDc, w.kCc, w.kDr, w.kRr, w.Dc, w.Cc, w.Dr, w.Rr, 0);
        if (TRACE) { ts.ck_widx = ts.widx; ts.ck_ridx = ts.ridx; }
        grow_no_max = false;
      }
      best_max = off_max;
      y_drop_iter = 0;
    }

    if (XDROP) {
      if (off_max < best_max - P.x_drop) {
        if (x_drop_iter < kXDropIter - 1) x_drop_iter++;
        else break;
      } else {
        x_drop_iter = 0;
      }
    }

    if (si + B > qlen && sj + B > rlen) break;
    if (sj + B > rlen) { si += kStep; dir = kDown; continue; }
    if (si + B > qlen) { sj += kStep; dir = kRight; continue; }

    const int next_size = B * 2;
    if (next_size <= max_size) {
      if (y_drop_iter > (B / kStep) - 1 || grow_no_max) {
        prev_size = B; B = next_size; dir = kGrow;
        si = i_ckpt; sj = j_ckpt; off = off_ckpt;
        copy4(prev_size, w.Dc, w.Cc, w.Dr, w.Rr, w.kDc, w.kCc, w.kDr, w.kRr, 0);
        if (TRACE) { ts.widx = ts.ck_widx; ts.ridx = ts.ck_ridx; }
        y_drop_iter = 0;
        continue;
      }
    }

    if (B > min_size && y_drop_iter == 0) {
      const int sm = shrink_max(w.Dc, w.Dr, B);
      if (sm >= mx) {
        prev_dir = kGrow;
        B /= 2;
        copy4(B, w.Dc, w.Cc, w.Dr, w.Rr, w.Dc, w.Cc, w.Dr, w.Rr, B);
        si += (uint32_t)B; sj += (uint32_t)B;
        i_ckpt = si; j_ckpt = sj; off_ckpt = off;
        copy4(B, w.kDc, w.kCc, w.kDr, w.kRr, w.Dc, w.Cc, w.Dr, w.Rr, 0);
        border_maxes(w.Dc, w.Dr, right_max, down_max);
        if (TRACE) { ts.ck_widx = ts.widx; ts.ck_ridx = ts.ridx; }
        y_drop_iter = 0;
      }
    }

    if (down_max > right_max) { si += kStep; dir = kDown; }
    else { sj += kStep; dir = kRight; }
  }

  // result (scan_block.rs:567-592)
  wp::syncwarp();
  DevResult res;
  res.status = ts.overflow ? (uint32_t)kTraceOverflow : (uint32_t)kOk;
  res.cells = cells; res.steps = steps; res.cigar_n = 0; res.cigar_off = 0;
  if (XDROP) {
    res.score = best_max; res.query_idx = best_i; res.reference_idx = best_j;
  } else {
    int sv;
    if (dir == kRight || dir == kGrow) sv = (int)w.Dc[qlen - si];
    else sv = (int)w.Dr[rlen - sj];
    res.score = off + sv - kZero; res.query_idx = qlen; res.reference_idx = rlen;
  }

  res.rect_n = ts.ridx; res.warp = warp_global;
  // traceback -> CIGAR runs (scan_block.rs:1482-1672; cigar.rs:71-79)
  if (TRACE && !ts.overflow && P.cigar_stream) {
    uint32_t* runs = P.run_scratch + (size_t)warp_global * P.runs_per_warp;
    emit_cigar(P, ts.words, ts.rects, ts.ridx, res.query_idx, res.reference_idx, q, r, P.cigar_eq != 0, runs, res);
  }
  if (lane == 0) P.out[pair] = res;
  wp::syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Warp main: carve shared memory, then pull pairs off the ticket counter until none are left.
// ---------------------------------------------------------------------------------------------
BA_HD size_t warp_smem_bytes(uint32_t max_size, bool ckpt_in_smem) {
  const size_t ms = max_size < 32 ? 32 : max_size;
  size_t b = 4 * ms * sizeof(int16_t) + 2 * 16 * sizeof(int16_t) + 4 * sizeof(int32_t);
  if (ckpt_in_smem) b += 4 * ms * sizeof(int16_t);
  if (max_size > 256) b += ms;  // ecarry
  return (b + 15) & ~(size_t)15;
}

template <int SCORING, int FLAGS>
BA_DEV void warp_main(const Params& P, unsigned char* smem, int warp_in_block, uint32_t warp_global) {
  const int lane = wp::lane_id();
  const size_t ms = P.max_size < 32 ? 32 : P.max_size;
  WarpMem w;
  w.mat = (const int8_t*)smem;                     // first 1 KB: matrix
  unsigned char* base = smem + 1024 + (size_t)warp_in_block * warp_smem_bytes(P.max_size, P.ckpt_in_smem != 0);
  int16_t* p16 = (int16_t*)base;
  w.Dc = p16; p16 += ms; w.Cc = p16; p16 += ms; w.Dr = p16; p16 += ms; w.Rr = p16; p16 += ms;
  w.t1 = p16; p16 += 16; w.t2 = p16; p16 += 16;
  w.misc = (int32_t*)p16; p16 += 8;
  if (P.ckpt_in_smem) {
    w.kDc = p16; p16 += ms; w.kCc = p16; p16 += ms; w.kDr = p16; p16 += ms; w.kRr = p16; p16 += ms;
  } else {
    int16_t* g = P.ckpt + (size_t)warp_global * 4 * ms;
    w.kDc = g; w.kCc = g + ms; w.kDr = g + 2 * ms; w.kRr = g + 3 * ms;
  }
  w.ecarry = (uint8_t*)p16;

  TraceState ts;
  ts.words = nullptr; ts.words_cap = 0; ts.rects = nullptr; ts.rects_cap = 0;
  ts.widx = 0; ts.ridx = 0; ts.ck_widx = 0; ts.ck_ridx = 0; ts.overflow = false;
  if (FLAGS & kTrace) {
    ts.words = P.trace_words + (size_t)warp_global * P.trace_words_per_warp;
    ts.words_cap = P.trace_words_per_warp;
    ts.rects = P.rects + (size_t)warp_global * P.rects_per_warp;
    ts.rects_cap = P.rects_per_warp;
  }

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = wp::atomic_add(P.ticket, 1u);
    t = (uint32_t)wp::shfl_idx((int)t, 0);
    if (t >= P.n_pairs) break;
    const uint32_t pair = P.order ? P.order[t] : t;
    align_pair<SCORING, FLAGS>(P, pair, w, ts, warp_global);
  }
}

// Traceback from an arbitrary end position of a pair that already ran (legacy block_cigar_* calls):
// the trace of pair `pair` is still in the arena of the warp that aligned it.
BA_DEV void warp_traceback(const Params& P, uint32_t pair, uint32_t qi, uint32_t rj, bool eq, DevResult* out1) {
  DevResult res = P.out[pair];
  const uint32_t wg = res.warp;
  const uint32_t* words = P.trace_words + (size_t)wg * P.trace_words_per_warp;
  const Rect* rects = P.rects + (size_t)wg * P.rects_per_warp;
  uint32_t* runs = P.run_scratch + (size_t)wg * P.runs_per_warp;
  const uint8_t* q = P.seq + P.q_off[pair];
  const uint8_t* r = P.profiles ? nullptr : P.seq + P.r_off[pair];
  res.status = (uint32_t)kOk;
  emit_cigar(P, words, rects, res.rect_n, qi, rj, q, r, eq, runs, res);
  if (wp::lane_id() == 0) *out1 = res;
}

}  // namespace ba
