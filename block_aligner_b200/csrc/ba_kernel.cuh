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

#ifdef BA_EMU
// coverage counters of the emulated build (tests assert that the packed path really ran)
namespace emu_stats { inline uint64_t pk_cells = 0, exact_cells = 0, fast_steps = 0, big_cells = 0, tb_staged = 0, tb_global = 0; }
#endif

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
  int16_t *kDc, *kCc, *kDr, *kRr;   // checkpoint copies of the current slot (global) [max_size]
  int32_t *misc;                    // [0..1]: A_old at the row above a chunk (double buffered)
  uint8_t *ecarry;                  // trace bit carried across chunks    [max_size] (only blocks > 256)
  const int8_t *mat;                // staged matrix (smem)
  const unsigned char* smem0;       // start of the CTA's shared memory (matrix + packed-path tables)
  uint32_t* fr;                     // packed path: bottom-row staging, 8 words per alignment group [64]
  uint32_t* ck;                     // packed fast phase: checkpoint borders in the packed layout, 64 words per group [512]
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
  int key_thr;                   // X_DROP, packed path: the argmax key is only needed if the block maximum exceeds this
  // extended modes (Block<_, _, LOCAL_START, FREE_QUERY_START_GAPS>; only read by EXT instantiations)
  bool local;                    // LOCAL_START: every cell is floored at relative_zero (scan_block.rs:1134-1136)
  bool fqs0;                     // FREE_QUERY_START_GAPS and this rectangle's vectors start at row 0 (:1130)
  uint32_t* tz;                  // zero-mask words (TRACE && LOCAL_START, scan_block.rs:1184-1187), same indexing as tw
  // FREE_QUERY_END_GAPS (scan_block.rs:333-368, 1194-1201): the block maximum is the running maximum of AVX lane
  // |q| % 16 only, and the argmax column is tracked for the vectors that reach row |q| or lie below it
  bool fqe; int fq_cls, fq_row0;  // lane class |q| % 16; first eligible row relative to the rectangle's top (= |q| - vec_base)
};

// profile context handed to the rectangle (scan_block.rs:612-783)
struct ProfArgs {
  const ProfileDev* p;
  const uint8_t* q;              // padded query
};

// Column loop is rolled (code size: the whole kernel must stay I-cache friendly); rows per lane R is
// 1 for rectangles up to 128 rows (swept in 32-row chunks) and 8 from 256 rows up (256-row chunks).
BA_HD int rect_rows_per_lane(int H) { return H >= 256 ? 8 : 1; }

template <int SCORING, bool RIGHT, int R, bool TRACE, bool XDROP, bool EXT>
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
  const int lane_rg = lane * R * ge;

  const int nchunks = (a.H + CH - 1) / CH;
  const bool multi = nchunks > 1;
  const int ngroups = a.W >> 3;
  const bool origin = (a.vec_base == 0 && a.col_base == 0);  // scan_block.rs:1130

  int fq_m = 0, fq_j = 0;      // D_max lane starts at MIN = 0, D_argmax_j at 0
  for (int ch = 0; ch < nchunks; ch++) {
    const int v0 = ch * CH + lane * R;
    const bool act = v0 < a.H;
    const int rows_here = wp::imin(CH, a.H - ch * CH);
    const int last_lane = rows_here / R - 1;

    int D10[R], C10[R], rtok[R];
    int p_openC[R], p_openR[R], p_closeR[R];   // profile-down per-row data
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
    if (multi) {
      // value of the old border at the row just above this chunk = diagonal input of column 0
      if (ch > 0) diag_carry = w.misc[ch & 1];
      if (lane == 31) w.misc[(ch + 1) & 1] = D10[R - 1];
      wp::syncwarp();
    }

    // per-row tracker: max of value * 16384 + (column + 1); 0 = "no cell >= 0" (D_max starts at MIN = 0)
    int trk[R];
    unsigned twd[R], zwd[R];
#pragma unroll
    for (int k = 0; k < R; k++) { trk[k] = 0; twd[k] = 0; zwd[k] = 0; }

#pragma unroll (R >= 8 ? 2 : 1)
    for (int cidx = 0; cidx < a.ncols; cidx++) {
      const int c4 = (cidx & 7) * 4;
      int Ttop = 0, Dtop = 0, etop = 0;
      if (multi) {
        if (ch > 0) {   // bottom row of the chunk above, written there for exactly this purpose
          Ttop = (int)a.OR_[cidx];
          Dtop = (int)a.OD[cidx];
          if (TRACE) etop = (int)w.ecarry[cidx];
        }
        wp::syncwarp();
      }
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
      diag_carry = Dtop;

      int dd[R], xx[R], c11[R], c11o[R], c11e[R], tt[R];
#pragma unroll
      for (int k = 0; k < R; k++) {
        int s;
        if (!PROF) s = sc.score(ctok, rtok[k]);
        else if (RIGHT) s = (int)wp::ldg(prow_col + rtok[k]);
        else s = (int)wp::ldg(p_row[k] + ctok);
        int d00 = (k == 0) ? up : D10[k - 1];
        if (ch == 0 && k == 0 && lane == 0 &&
            ((cidx == 0 && origin && !(EXT && a.local)) || (EXT && a.fqs0))) { d00 = a.rz; s = 0; }   // scan_block.rs:1130-1132
        const int oc = PROF ? (RIGHT ? openC_col : p_openC[k]) : go;
        c11o[k] = sat_add_lo(D10[k], oc);
        c11[k] = wp::viaddmax(C10[k], ge, c11o[k]);
        c11e[k] = c11[k];
        if (PROF && RIGHT) c11e[k] = sat_add(c11[k], closeC_col);   // C11_end (scan_block.rs:694)
        dd[k] = wp::imin(wp::viaddmax(d00, s, c11e[k]), kI16Max);
        if (EXT && a.local) dd[k] = wp::imax(dd[k], a.rz);
        const int orr = PROF ? (RIGHT ? openR_col : p_openR[k]) : open_r;
        xx[k] = sat_add_lo(dd[k], orr);
        tt[k] = (k == 0) ? xx[0] : wp::viaddmax(tt[k - 1], ge, xx[k]);
      }
      // ---- cross-lane max-plus scan of the lane aggregates ----
      int inc = tt[R - 1];
#pragma unroll
      for (int s = 0; s < 5; s++) {
        const int u = wp::shfl_up(inc, 1 << s);
        inc = wp::viaddmax(u, (R << s) * ge, inc);
      }
      int ex = wp::shfl_up(inc, 1);
      if (lane == 0) ex = kNegBig;
      const int cin = wp::viaddmax(Ttop, lane_rg, ex);

      int Dn[R], Tn[R];
      unsigned ebits = 0;
#pragma unroll
      for (int k = 0; k < R; k++) {
        Tn[k] = wp::viaddmax(cin, kg[k], tt[k]);
        if (TRACE || (PROF && !RIGHT)) {
          const int Rv = wp::imax(Tn[k], ph[k]);
          int Rend = Rv;
          if (PROF && !RIGHT) Rend = sat_add(Rv, p_closeR[k]);
          Dn[k] = wp::imax(dd[k], Rend);
          if (TRACE) {
            unsigned nib = (Dn[k] == c11e[k] ? 1u : 0u) | (Dn[k] == Rend ? 2u : 0u);
            nib |= (c11[k] == c11o[k] ? 4u : 0u);
            ebits |= (Rv == xx[k] ? 1u : 0u) << k;
            twd[k] |= nib << c4;
            if (EXT && a.local) zwd[k] |= (Dn[k] == a.rz ? 1u : 0u) << (cidx & 7);
          }
        } else {
          Dn[k] = wp::vimax3(dd[k], Tn[k], ph[k]);
        }
        if (XDROP) trk[k] = wp::imax(trk[k], Dn[k] * 16384 + (cidx + 1));
        else trk[0] = wp::imax(trk[0], act ? Dn[k] : 0);
      }
      if (EXT && a.fqe) {
        // The reference visits the cells of one AVX lane column by column, top vector first, and records the column
        // whenever an eligible cell is >= everything that lane has seen so far (D_max == D11 after the update).
        // Rows of one class sit in the same register slot (16 = 0 mod R) of every second lane.
        int vloc = Dn[0];
#pragma unroll
        for (int k = 1; k < R; k++) if ((a.fq_cls & (R - 1)) == k) vloc = Dn[k];
        for (int r0 = a.fq_cls; r0 < rows_here; r0 += 16) {
          const int v = wp::shfl_idx(vloc, r0 / R);
          if (v >= fq_m) { fq_m = v; if (ch * CH + r0 >= a.fq_row0) fq_j = cidx; }
        }
      }
      if (TRACE) {
        // "R gap at this row was opened at the row above" = e of the row above (scan_block.rs:1179-1181)
        const unsigned eb = wp::ballot(((ebits >> (R - 1)) & 1u) != 0);
        const unsigned above = lane == 0 ? (unsigned)etop : ((eb >> (lane - 1)) & 1u);
#pragma unroll
        for (int k = 0; k < R; k++) {
          const unsigned b3 = (k == 0) ? above : ((ebits >> (k - 1)) & 1u);
          twd[k] |= (b3 << 3) << c4;
        }
        if (multi && lane == last_lane) w.ecarry[cidx] = (uint8_t)((ebits >> (R - 1)) & 1u);
        if ((cidx & 7) == 7 || cidx == a.ncols - 1) {
          const int cgi = cidx >> 3;
#pragma unroll
          for (int k = 0; k < R; k++) {
            a.tw[(((size_t)ch * ngroups + cgi) * R + k) * 32 + lane] = twd[k];
            twd[k] = 0;
            if (EXT && a.local) { a.tz[(((size_t)ch * ngroups + cgi) * R + k) * 32 + lane] = zwd[k]; zwd[k] = 0; }
          }
        }
      }
      // ---- bottom row of the rectangle (or of this chunk) ----
      if (lane == last_lane) {
        a.OD[cidx] = (int16_t)Dn[R - 1];
        a.OR_[cidx] = (int16_t)Tn[R - 1];
      }
#pragma unroll
      for (int k = 0; k < R; k++) { D10[k] = Dn[k]; C10[k] = c11[k]; }
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
          const int v = trk[k] >> 14;
          const unsigned c1 = (unsigned)(trk[k] & 16383);
          const unsigned cls = (unsigned)((lane * R + k) & 15);
          const unsigned key = ((15u - cls) << kKeyClsShift) | (c1 << kKeyColShift) | (unsigned)(v0 + k);
          if (c1 != 0 && (v > bv || (v == bv && key > bkey))) { bv = v; bkey = key; }
        }
      }
    } else {
      bv = wp::imax(bv, trk[0]);
    }
    if (multi) wp::syncwarp();
  }
  if (EXT && a.fqe) { bv = fq_m; bkey = (unsigned)fq_j; }
  wp::syncwarp();
}

// RIGHT only changes behaviour for profiles (sequence-sequence right/down differ only in which
// sequence plays which role), so sequence scorers always instantiate RIGHT = true.
template <int SCORING, bool RIGHT, bool TRACE, bool XDROP, bool EXT>
BA_DEV void place_rect(const SeqScorer<(SCORING == kProfile ? kAA : SCORING)>& sc, const ProfArgs& pa,
                       const RectArgs& a, const WarpMem& w, int& bv, unsigned& bkey) {
  constexpr bool RT = (SCORING == kProfile) ? RIGHT : true;
  bv = 0;                // D_max starts at MIN = 0 (scan_block.rs:1101)
  bkey = 15u << kKeyClsShift;      // "no cell": AVX lane 0 with argmax (0, 0)
  if (a.W == 0 || a.H == 0) return;   // scan_block.rs:1105-1107
  // FREE_QUERY_END_GAPS needs the whole rectangle in one chunk (cell order matters): 8 rows per lane up to 256 rows
  if (a.H >= 256 || (EXT && a.fqe)) place_rect_r<SCORING, RT, 8, TRACE, XDROP, EXT>(sc, pa, a, w, bv, bkey);
  else place_rect_r<SCORING, RT, 1, TRACE, XDROP, EXT>(sc, pa, a, w, bv, bkey);
}

}  // namespace ba
#include "ba_packed.cuh"
namespace ba {

// ---------------------------------------------------------------------------------------------
// Border helpers
// ---------------------------------------------------------------------------------------------
#ifndef BA_SHIFT_VEC
#define BA_SHIFT_VEC 1
#endif
// shift_and_offset (scan_block.rs:1040-1061): slide by STEP, re-offset the kept part, append the
// 8 fresh values from temp (not re-offset). Returns old buf1[STEP-1] + off_add.
BA_DEV uint32_t sat_add_pair(uint32_t p, int off_add) {   // both halfwords: _mm256_adds_epi16 with a scalar
  return wp::h_pack(sat_add(wp::h_lo(p), off_add), sat_add(wp::h_hi(p), off_add));
}
BA_DEV int shift_and_offset(int B, int16_t* b1, int16_t* b2, const int16_t* t1, const int16_t* t2, int off_add) {
  const int lane = wp::lane_id();
  const int corner = sat_add((int)b1[kStep - 1], off_add);
  wp::syncwarp();
#if BA_SHIFT_VEC
  // eight entries (16 bytes) per lane and pass: chunk ch of the result is chunk ch + 1 of the input, re-offset; the last
  // chunk is the fresh one. A pass reads chunks [32 it + 1, 32 it + 33) and then writes [32 it, 32 it + 32): nothing is
  // overwritten before it has been read.
  const int nch = B / kStep;
  for (int it = 0; it * 32 < nch; it++) {
    const int ch = it * 32 + lane;
    uint4 v1 = make_uint4(0u, 0u, 0u, 0u), v2 = v1;
    if (ch < nch) {
      if (ch < nch - 1) {
        v1 = *(const uint4*)(b1 + kStep * (ch + 1)); v2 = *(const uint4*)(b2 + kStep * (ch + 1));
        v1.x = sat_add_pair(v1.x, off_add); v1.y = sat_add_pair(v1.y, off_add); v1.z = sat_add_pair(v1.z, off_add); v1.w = sat_add_pair(v1.w, off_add);
        v2.x = sat_add_pair(v2.x, off_add); v2.y = sat_add_pair(v2.y, off_add); v2.z = sat_add_pair(v2.z, off_add); v2.w = sat_add_pair(v2.w, off_add);
      } else {
        v1 = *(const uint4*)t1; v2 = *(const uint4*)t2;
      }
    }
    wp::syncwarp();
    if (ch < nch) { *(uint4*)(b1 + kStep * ch) = v1; *(uint4*)(b2 + kStep * ch) = v2; }
    wp::syncwarp();
  }
#else
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
#endif
  return corner;
}

BA_DEV void copy4(int n, int16_t* d0, int16_t* d1, int16_t* d2, int16_t* d3,
                  const int16_t* s0, const int16_t* s1, const int16_t* s2, const int16_t* s3, int src_off) {
  const int lane = wp::lane_id();
  wp::syncwarp();
  // n and src_off are multiples of 8 entries (block sizes are powers of two >= 16) and the arrays start on 16-byte
  // boundaries: eight entries per lane and pass. Source and destination may be the same arrays with src_off = n (shrink).
  for (int base = 0; base < n; base += 8 * 32) {
    const int idx = base + 8 * lane;
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a, c = a, d = a;
    if (idx < n) { a = *(const uint4*)(s0 + idx + src_off); b = *(const uint4*)(s1 + idx + src_off); c = *(const uint4*)(s2 + idx + src_off); d = *(const uint4*)(s3 + idx + src_off); }
    wp::syncwarp();
    if (idx < n) { *(uint4*)(d0 + idx) = a; *(uint4*)(d1 + idx) = b; *(uint4*)(d2 + idx) = c; *(uint4*)(d3 + idx) = d; }
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
// Per-alignment state. Everything the reference keeps in locals of align_core (scan_block.rs:
// 102-128) plus the trace stack indices, so that an alignment can be parked and resumed: the
// block-32 shift steps run four alignments per warp out of registers (fast phase, below), everything
// else runs one alignment per warp out of shared memory (generic phase).
// ---------------------------------------------------------------------------------------------
struct AlnState {
  uint32_t pair, qlen, rlen;
  uint32_t si, sj;
  int B, prev_size, dir, prev_dir;
  int off, off_max, best_max;
  uint32_t best_i, best_j, i_ckpt, j_ckpt;
  int off_ckpt, y_drop_iter, x_drop_iter, D_corner;
  uint32_t cells_lo, cells_hi, steps;
  uint32_t widx, ridx, ck_widx, ck_ridx, overflow;   // trace stack (scan_block.rs:1351-1354)
  uint32_t ck_pk;                                    // the newest checkpoint borders are in WarpMem::ck (packed fast phase)
};

// per-slot scratch in global memory (slot = one alignment in flight)
struct SlotMem {
  int16_t *kDc, *kCc, *kDr, *kRr;     // checkpoint borders (scan_block.rs:1262-1265)
  uint32_t* words; uint64_t words_cap;
  uint32_t* zwords;                   // zero masks (TRACE && LOCAL_START only)
  Rect* rects; uint32_t rects_cap;
};

BA_DEV void add_cells(AlnState& st, uint32_t n) {
  const uint32_t lo = st.cells_lo + n;
  st.cells_hi += (lo < st.cells_lo) ? 1u : 0u;
  st.cells_lo = lo;
}

// layout: 0 = rows-per-lane from the height, 1 = forced 8 rows per lane (FREE_QUERY_END_GAPS), 3 = packed path
// (one word per lane and column, word index = column * H / 8 + lane; ba_packed.cuh)
BA_HD uint64_t rect_words(int H, int W, int layout = 0) {
  if (layout == 3 || layout == 2) return (uint64_t)(H >> 3) * (uint64_t)W;     // packed path (2: taller than 256 rows, in 256-row chunks)
  const int R = layout == 1 ? 8 : rect_rows_per_lane(H);
  const int CH = 32 * R;
  const int nch = (H + CH - 1) / CH;
  return (uint64_t)nch * (uint64_t)(W >> 3) * R * 32;
}
// Trace::add_block (scan_block.rs:1428-1443); returns where this rectangle's words go.
// The word layout of the rectangle is recorded in bits 1-2 of Rect::right so that the walk-back indexes the words
// the same way. A rectangle that does not fit the slot's own arena is placed in the batch-wide overflow pool
// (bit 3 of Rect::right, word_off in units of 16 words): slot arenas are sized for the common path, and the few
// alignments that sit at large block sizes for long borrow from the pool instead of being re-run.
// Called by all 32 lanes of a warp that services one alignment (generic phase); `writer`: the lane that stores the
// record.
constexpr uint32_t kRectPool = 8u;
BA_DEV uint32_t* trace_push(const Params& P, AlnState& st, const SlotMem& sm, uint32_t row, uint32_t col, int W, int H, bool right,
                            bool writer, int layout = 0, bool want = true, int GW = 32) {
  const uint64_t need = rect_words(H, W, layout);
  bool over = false, pool = false;
  if (want) {
    if (st.ridx >= sm.rects_cap) over = true;
    else if ((uint64_t)st.widx + need > sm.words_cap || ((uint64_t)st.widx + need) >> 32) pool = true;
  }
  uint32_t pbase = 0;
  if (wp::ballot(pool) != 0u) {
    const uint32_t units = (uint32_t)((need + 15) >> 4);
    if (pool && writer && P.trace_pool) pbase = wp::atomic_add(P.trace_pool_cursor, units);
    pbase = (uint32_t)wp::shfl_idx_w((int)pbase, 0, GW);
    if (pool && (!P.trace_pool || (uint64_t)pbase + units > P.trace_pool_units)) { over = true; pool = false; }
  }
  if (!want) return sm.words;
  if (over) { st.overflow = 1u; return sm.words; }
  Rect r; r.row = row; r.col = col; r.h = (uint16_t)H; r.w = (uint16_t)W; r.right = (right ? 1u : 0u) | ((uint32_t)layout << 1);
  uint32_t* p;
  if (pool) { r.right |= kRectPool; r.word_off = pbase; p = P.trace_pool + (size_t)pbase * 16; }
  else { r.word_off = st.widx; p = sm.words + st.widx; st.widx += (uint32_t)need; }
  if (writer) sm.rects[st.ridx] = r;
  st.ridx += 1;
  return p;
}
// Fast-phase variant: the slot's own arena only, no pool (the atomic + shuffle of the pool path in the middle of the
// step cost 1000 bytes of extra register spills per thread). A group whose arena cannot take another step is parked
// first (trace_room) and finishes in the generic phase, which can use the pool.
BA_DEV uint32_t* trace_push_local(AlnState& st, const SlotMem& sm, uint32_t row, uint32_t col, int W, int H, bool right, bool writer, int layout) {
  const uint64_t need = rect_words(H, W, layout);
  if (st.ridx >= sm.rects_cap || (uint64_t)st.widx + need > sm.words_cap || ((uint64_t)st.widx + need) >> 32) {
    st.overflow = 1u;
    return sm.words;
  }
  if (writer) {
    Rect r; r.row = row; r.col = col; r.h = (uint16_t)H; r.w = (uint16_t)W; r.right = (right ? 1u : 0u) | ((uint32_t)layout << 1); r.word_off = st.widx;
    sm.rects[st.ridx] = r;
  }
  uint32_t* p = sm.words + st.widx;
  st.widx += (uint32_t)need;
  st.ridx += 1;
  return p;
}
// room for one more fast-phase step (B rows x 8 columns, layout 3: B words, one record) in the slot's own arena
BA_DEV bool trace_room(const AlnState& st, uint64_t words_cap, uint32_t rects_cap, int B) {
  return (uint64_t)st.widx + (uint64_t)B <= words_cap && st.ridx < rects_cap;
}
BA_DEV const uint32_t* rect_words_ptr(const uint32_t* words, const uint32_t* pool, const Rect& rc) {
  return (rc.right & kRectPool) ? pool + (size_t)rc.word_off * 16 : words + rc.word_off;
}

// ---------------------------------------------------------------------------------------------
// Traceback (Trace::cigar_core, scan_block.rs:1482-1672). Lane 0 walks from (i, j) back to the
// origin through the stack of rectangles; runs are produced reversed and run-length merged
// exactly like Cigar::add (cigar.rs:71-79), then written forward into the output stream.
// The walk is one dependent chain (trace word -> step -> next cell), so its speed is the latency of
// that chain: the other lanes keep the trace words of the rectangles the walk reaches next in L1
// (the words were written tens of thousands of steps ago and live in DRAM by now), and the
// per-cell decision is one lookup in a 128-entry table instead of a branch tree.
// ---------------------------------------------------------------------------------------------
struct WalkState {
  uint32_t i, j, ridx, table, cur_op, cur_len, nruns, bad, stop;
};

// lane 0: walk inside rectangle rc until the path leaves it
BA_DEV void traceback_rect(const uint8_t* lut, const uint32_t* words, const uint32_t* pool, const uint32_t* zwords, bool fqs, const Rect& rc,
                           const uint8_t* q, const uint8_t* r, bool eq, uint32_t* runs, uint32_t cap, WalkState& s) {
  const int H = rc.h, W = rc.w;
  const bool rc_right = (rc.right & 1u) != 0;
  const uint32_t layout = (rc.right >> 1) & 3u;
  const int R = layout == 1u ? 8 : rect_rows_per_lane(H);
  const int CH = 32 * R;
  const int ngroups = W >> 3;
  const uint32_t* tw = rect_words_ptr(words, pool, rc);
  const uint32_t lbase = rc_right ? 64u : 0u;
  const uint32_t hh = (uint32_t)(H >> 1), G = (uint32_t)(H >> 3);
  uint32_t i = s.i, j = s.j, table = s.table;
  while (i >= rc.row && j >= rc.col && (i > 0 || j > 0)) {
    // FREE_QUERY_START_GAPS: stop on row 0, which always lies in right rectangles (scan_block.rs:1597-1600)
    if (fqs && rc_right && i == 0) { s.stop = 1u; break; }
    const uint32_t v = rc_right ? i - rc.row : j - rc.col;
    const uint32_t c = rc_right ? j - rc.col : i - rc.row;
    uint32_t nib;
    if (layout >= 2u) {     // packed path: word (column, lane in group), nibble (row mod 4) of the half-block's 16 bits;
      // layout 2: the rectangle is a stack of 256-row chunks of 32 lanes each
      const uint32_t chn = layout == 2u ? (v >> 8) : 0u, v8 = layout == 2u ? (v & 255u) : v, h2 = layout == 2u ? 128u : hh;
      const uint32_t half = v8 >= h2 ? 1u : 0u, vv = v8 - half * h2;
      nib = (tw[(size_t)c * G + chn * 32u + (vv >> 2)] >> (16u * half + 4u * (vv & 3u))) & 15u;
    } else {
      const uint32_t ch = v / CH, ln = (v % CH) / R, k = v % R;
      const size_t widx = (((size_t)ch * ngroups + (c >> 3)) * R + k) * 32 + ln;
      // LOCAL_START: the alignment starts at a cell equal to relative_zero (scan_block.rs:1606-1612)
      if (zwords && table == 0 && ((zwords[rc.word_off + widx] >> (c & 7)) & 1u)) { s.stop = 1u; break; }
      nib = (tw[widx] >> (4 * (c & 7))) & 15u;
    }
    const uint32_t e = lut[lbase | (table << 4) | nib];
    uint32_t op = e & 7u;
    const uint32_t di = (e >> 3) & 1u, dj = (e >> 4) & 1u;
    if (op == 1u && eq) op = (q[i] == r[j]) ? 2u : 3u;   // scan_block.rs:1620-1628
    if ((di && i == 0) || (dj && j == 0)) { s.bad = 1u; break; }
    i -= di; j -= dj; table = e >> 5;
    if (op == s.cur_op) s.cur_len++;
    else {
      if (s.cur_len) { if (s.nruns < cap) runs[s.nruns] = (s.cur_len << 4) | s.cur_op; s.nruns++; }
      s.cur_op = op; s.cur_len = 1;
    }
  }
  s.i = i; s.j = j; s.table = table;
}

// Whole warp. Lane 0 walks; the other lanes touch the trace words of the rectangle kTbAhead records further down
// the stack (the next ones the walk will usually enter), so that they are in L1 when lane 0 gets there.
constexpr uint32_t kTbAhead = 8;
BA_DEV void traceback_walk(const uint8_t* lut, const uint32_t* words, const uint32_t* pool, const uint32_t* zwords, bool fqs, const Rect* rects,
                           uint32_t ridx, uint32_t i, uint32_t j, const uint8_t* q, const uint8_t* r, bool eq, uint32_t* runs,
                           uint32_t cap, uint32_t& nruns, uint32_t& bad) {
  const int lane = wp::lane_id();
  WalkState s;
  s.i = i; s.j = j; s.ridx = ridx; s.table = 0; s.cur_op = 0; s.cur_len = 0; s.nruns = 0; s.bad = 0; s.stop = 0;
  for (;;) {
    uint32_t go = 0;
    Rect rc;
    rc.row = 0; rc.col = 0; rc.h = 0; rc.w = 0; rc.right = 0; rc.word_off = 0;
    if (lane == 0 && (s.i > 0 || s.j > 0) && !s.bad && !s.stop) {
      for (;;) {   // find the newest rectangle containing (i, j) (scan_block.rs:1578-1590)
        if (s.ridx == 0) { s.bad = 1u; break; }
        s.ridx--;
        rc = rects[s.ridx];
        if (s.i >= rc.row && s.j >= rc.col) break;
      }
      go = s.bad ? 0u : 1u;
    }
    go = (uint32_t)wp::shfl_idx((int)go, 0);
    if (!go) break;
    const uint32_t cur = (uint32_t)wp::shfl_idx((int)s.ridx, 0);
    if (lane == 0) {
      traceback_rect(lut, words, pool, zwords, fqs, rc, q, r, eq, runs, cap, s);
    } else if (lane <= 4 && cur >= kTbAhead) {
      const Rect pr = rects[cur - kTbAhead];
      const uint64_t nw = (uint64_t)(pr.h >> 3) * pr.w;     // layout 3 shift rectangles: h words, 32 per 128-byte line
      if (((pr.right >> 1) & 3u) == 3u && nw <= 128u && (uint64_t)(lane - 1) * 32u < nw) wp::touch(rect_words_ptr(words, pool, pr) + (lane - 1) * 32);
    }
    wp::syncwarp();
  }
  if (lane == 0 && s.cur_len) { if (s.nruns < cap) runs[s.nruns] = (s.cur_len << 4) | s.cur_op; s.nruns++; }
  nruns = s.nruns; bad = s.bad;
}

BA_DEV void emit_cigar(const Params& P, const uint8_t* lut, const uint32_t* words, const uint32_t* zwords, bool fqs, const Rect* rects,
                       uint32_t ridx, uint32_t qi, uint32_t rj, const uint8_t* q, const uint8_t* r, bool eq, uint32_t* runs,
                       DevResult& res) {
  const int lane = wp::lane_id();
  uint32_t nruns = 0, bad = 0;
  traceback_walk(lut, words, P.trace_pool, zwords, fqs, rects, ridx, qi, rj, q, r, eq, runs, P.runs_per_warp, nruns, bad);
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
// Generic phase: one alignment per warp, borders in shared memory (scan_block.rs:94-595)
// ---------------------------------------------------------------------------------------------
enum { kRunDone = 0, kRunFast = 1 };

template <int SCORING, int FLAGS>
BA_DEV void init_alignment(const Params& P, AlnState& st, uint32_t pair, const WarpMem& w) {
  const int lane = wp::lane_id();
  st.pair = pair;
  st.qlen = P.q_len[pair];
  st.rlen = (SCORING == kProfile) ? P.profiles[pair].len : P.r_len[pair];
  // Allocated::clear (scan_block.rs:1322-1339): every border starts at MIN = 0
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  for (int idx = 8 * lane; idx < (int)P.max_size; idx += 8 * 32) {      // max_size is a multiple of 16 entries
    *(uint4*)(w.Dc + idx) = z4; *(uint4*)(w.Cc + idx) = z4; *(uint4*)(w.Dr + idx) = z4; *(uint4*)(w.Rr + idx) = z4;
    *(uint4*)(w.kDc + idx) = z4; *(uint4*)(w.kCc + idx) = z4; *(uint4*)(w.kDr + idx) = z4; *(uint4*)(w.kRr + idx) = z4;
  }
  if (lane < 16) { w.t1[lane] = 0; w.t2[lane] = 0; }
  wp::syncwarp();
  st.si = 0; st.sj = 0;
  st.B = (int)P.min_size; st.prev_size = 0; st.dir = kGrow; st.prev_dir = kGrow;
  st.off = 0; st.off_max = 0; st.best_max = 0; st.best_i = 0; st.best_j = 0;
  st.i_ckpt = 0; st.j_ckpt = 0; st.off_ckpt = 0;
  st.y_drop_iter = 0; st.x_drop_iter = 0; st.D_corner = 0;
  st.cells_lo = 0; st.cells_hi = 0; st.steps = 0;
  st.widx = 0; st.ridx = 0; st.ck_widx = 0; st.ck_ridx = 0; st.overflow = 0; st.ck_pk = 0;
}

// grow transition (scan_block.rs:477-500): back to the checkpoint with a doubled block
BA_DEV void apply_grow(AlnState& st, const WarpMem& w, bool trace) {
  st.prev_size = st.B; st.B = st.B * 2; st.dir = kGrow;
  st.si = st.i_ckpt; st.sj = st.j_ckpt; st.off = st.off_ckpt;
  copy4(st.prev_size, w.Dc, w.Cc, w.Dr, w.Rr, w.kDc, w.kCc, w.kDr, w.kRr, 0);
  if (trace) { st.widx = st.ck_widx; st.ridx = st.ck_ridx; }
  st.y_drop_iter = 0;
}

// Is the pending step (st.dir already chosen, st.si/sj already moved) a plain block-32 shift that the
// fast phase can execute? (no early break: scan_block.rs:1216-1224)
template <int SCORING, bool XDROP>
BA_DEV bool shift_eligible(const AlnState& st, int FB) {
  if (FB == 0) return false;
  if (st.B != FB || st.dir == kGrow) return false;
  if (!XDROP) {
    const uint32_t vec_base = st.dir == kRight ? st.si : st.sj, vec_len = st.dir == kRight ? st.qlen : st.rlen;
    const uint32_t col_base = (st.dir == kRight ? st.sj : st.si) + (uint32_t)(FB - kStep), col_len = st.dir == kRight ? st.rlen : st.qlen;
    if (vec_base + (uint32_t)FB > vec_len && col_base + 7 > col_len) return false;
  }
  return true;
}
template <int SCORING, bool XDROP>
BA_DEV bool fast_eligible(const Params& P, const AlnState& st) {
  return shift_eligible<SCORING, XDROP>(st, (int)P.fast_block);     // 0 (fast phase off), 32 or 64 == min block size
}

// The packed fast phase keeps its newest checkpoint in shared memory (pk_fast_step); the generic phase reads and
// writes the slot's checkpoint in global memory in the plain layout. Called when a group parks.
template <int LGT>
BA_DEV void pk_ckpt_flush(const WarpMem& w, const SlotMem& sm, int g, bool mine) {
  constexpr int G = 1 << LGT;
  const int lg = wp::lane_id() & (G - 1);
  if (mine) {
    const uint4* ck = (const uint4*)w.ck + (g * 4) * G + lg;
    int16_t* dst[4] = {sm.kDc, sm.kCc, sm.kDr, sm.kRr};
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const uint4 v = ck[a * G];
      const uint32_t r[4] = {v.x, v.y, v.z, v.w};
      pk_store4(dst[a], lg, G, r);
    }
  }
  wp::syncwarp();
}

// Are the borders (shared memory, generic phase) inside the range the next packed shift step needs?
BA_DEV bool pk_borders_ok(const Params& P, const AlnState& st, const WarpMem& w) {
  const int lane = wp::lane_id();
  int GL, GH;
  pk_bounds(kStep, P.gap_open, P.gap_extend, P.pk_smax, GL, GH);
  if (P.scoring == kProfile) { GL = kPkpGL; GH = kPkpGH; }
  const int noa = clamp16(st.off - st.off_max);
  const int lo_b = wp::imax(GL - noa, kI16Min), hi_b = wp::imin(GH - noa, kI16Max);
  bool ok = lo_b <= hi_b && noa >= -1024 && noa <= 1024;
  for (int idx = lane; idx < st.B; idx += 32) {
    const int a = w.Dc[idx], b = w.Cc[idx], c = w.Dr[idx], d = w.Rr[idx];
    const int mn = wp::imin(wp::imin(a, b), wp::imin(c, d)), mx = wp::imax(wp::imax(a, b), wp::imax(c, d));
    ok = ok && mn >= lo_b && mx <= hi_b;
  }
  if (st.prev_dir != kGrow && st.prev_dir != st.dir) {
    const int cv = sat_add(st.D_corner, noa);
    ok = ok && cv >= 0 && cv <= GH;
  }
  return wp::ballot(!ok) == 0u;
}

// BA_PK_ORIGIN = 1: the first block of an alignment goes through the packed path (run_generic); 0: exact 32-bit path
#ifndef BA_PK_ORIGIN
#define BA_PK_ORIGIN 1
#endif
enum { kStFast = 0, kStNeedGeneric = 1, kStNeedGrow = 2, kStDone = 3, kStEmpty = 4, kStNeedShrink = 5 };
// cw (BA_CW_PREFETCH = 1 only): the eight column tokens of the group's next step, loaded while the previous step's
// epilogue runs -- the first use of that load is the largest single stall of the fast phase (3.7 % of the C2 kernel's
// samples, ncu r02). OFF: measured on B200 / C2 the two registers it keeps alive across the step push three others
// into local memory (6 LDL per step) and the kernel falls from 1195 to 1128 GCUPS.
#ifndef BA_CW_PREFETCH
#define BA_CW_PREFETCH 0
#endif
struct PkFast { uint32_t aD[4], aC[4], oD[4], oR[4]; uint2 cw; };

// load the column tokens of the pending shift step (st.dir / st.si / st.sj already describe it)
template <int SCORING, int LGT>
BA_DEV void pk_fast_prefetch(PkFast& f, const AlnState& st, const uint8_t* qp, const uint8_t* rp) {
  constexpr int B = 8 << LGT;
  const bool right = st.dir == kRight;
  if (SCORING == kProfile && right) return;     // a right step reads the profile's columns, not tokens
  const uint8_t* col = right ? rp : qp;
  f.cw = *(const uint2*)(col + (right ? st.sj : st.si) + (B - kStep));
}

template <int LGT>
BA_DEV void pk_fast_load(PkFast& f, const WarpMem& w, int dir, bool mine) {
  constexpr int G = 1 << LGT;
  const int lg = wp::lane_id() & (G - 1);
  const int16_t *ad = dir == kRight ? w.Dc : w.Dr, *ac = dir == kRight ? w.Cc : w.Rr;
  const int16_t *od = dir == kRight ? w.Dr : w.Dc, *orr = dir == kRight ? w.Rr : w.Cc;
  if (mine) { pk_load4(ad, lg, G, f.aD); pk_load4(ac, lg, G, f.aC); pk_load4(od, lg, G, f.oD); pk_load4(orr, lg, G, f.oR); }
}
template <int LGT>
BA_DEV void pk_fast_spill(const PkFast& f, const WarpMem& w, int dir, bool mine) {
  constexpr int G = 1 << LGT;
  const int lg = wp::lane_id() & (G - 1);
  int16_t *ad = dir == kRight ? w.Dc : w.Dr, *ac = dir == kRight ? w.Cc : w.Rr;
  int16_t *od = dir == kRight ? w.Dr : w.Dc, *orr = dir == kRight ? w.Rr : w.Cc;
  wp::syncwarp();
  if (mine) {
    pk_store4(ad, lg, G, f.aD); pk_store4(ac, lg, G, f.aC); pk_store4(od, lg, G, f.oD); pk_store4(orr, lg, G, f.oR);
    // temp_buf1/2 hold the 8 fresh values of the last shift = the last 8 entries of the orthogonal border
    if (lg >= G - 2) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        w.t1[4 * (lg - (G - 2)) + k] = (int16_t)wp::h_hi(f.oD[k]);
        w.t2[4 * (lg - (G - 2)) + k] = (int16_t)wp::h_hi(f.oR[k]);
      }
    }
  }
  wp::syncwarp();
}

template <int SCORING, int FLAGS, int LGT, bool BIG>
BA_DEV void pk_fast_step(const Params& P, const WarpMem& w, AlnState& st, PkFast& f, int& status,
                         const uint8_t* qp, const uint8_t* rp, uint32_t my_slot);
BA_DEV void bcast_state(AlnState& d, const AlnState& s, int src);

// Register-resident shift steps for a block above the minimum size (64 / 128 / 256 rows; the unrelated tails of X-drop
// alignments sit at the maximum size for dozens of steps): the alignment being serviced by the generic phase runs
// pk_fast_step with 1 << LGT lanes until a step needs something else. Returns the status that ended the loop; borders
// are back in shared memory, the newest checkpoint in the slot's global checkpoint.
#ifndef BA_BIG_LOOP_INLINE
#define BA_BIG_LOOP_FN BA_DEV_NOINLINE     // own register allocation: inlined into run_generic it made the whole kernel spill
#else
#define BA_BIG_LOOP_FN BA_DEV
#endif
template <int SCORING, int FLAGS, int LGT>
BA_BIG_LOOP_FN int big_loop(const Params& P, AlnState* stp, const WarpMem* wp_, uint32_t slot) {
  // by-value working copies: the caller's state and pointers stay in registers on its side of the call (only the
  // temporaries it passes are address-taken), and this function keeps its own copies in registers
  AlnState st = *stp;
  const WarpMem w = *wp_;
  constexpr int G = 1 << LGT;
  const int lane = wp::lane_id();
  const bool mine = lane < G;
  PkFast f;
#pragma unroll
  for (int k = 0; k < 4; k++) { f.aD[k] = 0; f.aC[k] = 0; f.oD[k] = 0; f.oR[k] = 0; }
  pk_fast_load<LGT>(f, w, st.dir, mine);
  const uint8_t* qp = P.seq + P.q_off[st.pair];
  const uint8_t* rp = P.seq + P.r_off[st.pair];
  f.cw = make_uint2(0u, 0u);
  if (BA_CW_PREFETCH) pk_fast_prefetch<SCORING, LGT>(f, st, qp, rp);
  int status = mine ? kStFast : kStEmpty;
  for (;;) {
    pk_fast_step<SCORING, FLAGS, LGT, true>(P, w, st, f, status, qp, rp, slot);
    if (wp::shfl_idx(status, 0) != kStFast) break;
  }
  status = wp::shfl_idx(status, 0);
  if (G < 32) bcast_state(st, st, 0);
  // the registers are laid out for the step executed last (= prev_dir)
  pk_fast_spill<LGT>(f, w, st.prev_dir, mine);
  *stp = st;
  return status;
}

// TALL: rectangles taller than 256 rows may take the packed path (place_rect_pk_tall). Only kernels for max block sizes
// >= 1024 (FM >= 32) carry that code: elsewhere it would be 9 KB that never (max <= 256) or hardly ever (512) runs.
template <int SCORING, int FLAGS, bool TALL = false>
BA_DEV int run_generic(const Params& P, AlnState& st, const WarpMem& w, const SlotMem& sm, uint32_t slot) {
  constexpr bool TRACE = (FLAGS & kTrace) != 0, XDROP = (FLAGS & kXDrop) != 0, EXT = (FLAGS & kExt) != 0;
  constexpr bool PROF = SCORING == kProfile;
  const bool m_local = EXT && (P.ext_flags & kLocalStart), m_fqs = EXT && (P.ext_flags & kFreeQueryStartGaps);
  const bool m_fqe = EXT && (P.ext_flags & kFreeQueryEndGaps);
  const int lane = wp::lane_id();
  const uint8_t* q = P.seq + P.q_off[st.pair];
  const uint8_t* r = nullptr;
  ProfArgs pa; pa.p = nullptr; pa.q = q;
  if (PROF) pa.p = P.profiles + st.pair;
  else r = P.seq + P.r_off[st.pair];
  const uint32_t qlen = st.qlen, rlen = st.rlen;

  SeqScorer<(SCORING == kProfile ? kAA : SCORING)> sc;
  sc.mat = w.mat; sc.vec = q; sc.col = r; sc.go = P.gap_open; sc.ge = P.gap_extend;
  sc.b_match = sc.b_mismatch = 0;
  if (SCORING == kByte) { sc.b_match = (int)P.matrix[0]; sc.b_mismatch = (int)P.matrix[1]; }
  const int min_size = (int)P.min_size, max_size = (int)P.max_size;

  bool resume_shrink = false;
  for (;;) {
    int bv = 0, gbv = 0; unsigned bkey = 15u << kKeyClsShift, gbkey = 15u << kKeyClsShift;
    int right_max = 0, down_max = 0, mx = 0;
    const int B = st.B;
    const uint32_t si = st.si, sj = st.sj;
    if (!resume_shrink) {
    if (fast_eligible<SCORING, XDROP>(P, st) && (!TRACE || trace_room(st, sm.words_cap, sm.rects_cap, st.B)) &&
        (!P.pk_fast || pk_borders_ok(P, st, w))) return kRunFast;
#ifdef BA_BIG_LOOP
    // Shift steps of a block above the minimum size in a register-resident loop (big_loop). OFF: measured on B200 / C2
    // (100 k pairs) the kernel fell from 1150 GCUPS to 913-1021 in every variant (all sizes / 256 only, inlined /
    // noinline): inlined, the extra copies of the step code push the kernel from 93 KB to 158 KB and 371 spill
    // instructions; as a separate function the ~60 live registers of the caller are saved around every call. The
    // instruction cache and the 128-register cap, not the instruction count of the tail steps, decide.
#ifndef BA_BIG_SIZES
#define BA_BIG_SIZES (64 | 128 | 256)     // block sizes served by big_loop (tuning: every size is one more copy of the step code)
#endif
    if (!PROF && !EXT && P.pk_fast && (B == 64 || B == 128 || B == 256) && (B & BA_BIG_SIZES) && B > (int)P.fast_block && !st.overflow &&
        shift_eligible<SCORING, XDROP>(st, B) && (!TRACE || trace_room(st, sm.words_cap, sm.rects_cap, B)) && pk_borders_ok(P, st, w)) {
      int bs;
      AlnState tst = st;       // temporaries for the call (see big_loop)
      WarpMem tw = w;
      if ((BA_BIG_SIZES & 256) && B == 256) bs = big_loop<SCORING, FLAGS, 5>(P, &tst, &tw, slot);
      else if ((BA_BIG_SIZES & 128) && B == 128) bs = big_loop<SCORING, FLAGS, 4>(P, &tst, &tw, slot);
      else if (BA_BIG_SIZES & 64) bs = big_loop<SCORING, FLAGS, 3>(P, &tst, &tw, slot);
      else bs = kStNeedGeneric;
      st = tst;
      if (bs == kStDone) return kRunDone;
      if (bs == kStNeedGrow) { apply_grow(st, w, TRACE); continue; }
      if (bs == kStNeedShrink) { resume_shrink = true; continue; }
      continue;    // kStNeedGeneric: the pending step fails one of the tests above and runs below
    }
#endif
    }
    if (resume_shrink) {
      // a big_loop step found a new best with the maximum in the corner: resume after the grow test (scan_block.rs:505)
      resume_shrink = false;
      mx = st.off_max - st.off + kZero;
      border_maxes(w.Dc, w.Dr, right_max, down_max);
    } else {
    const int prev_off = st.off;
    st.steps++;
    // One shift step computes one rectangle, a grow step two (down part, then right part). A single
    // place_rect call site serves all of them: the kernel has to stay small enough for the I-cache.
    int off_add = 0;
    if (st.dir == kGrow) st.D_corner = 0;
    else { st.off = st.off_max; off_add = clamp16(prev_off - st.off); }
    const int nparts = st.dir == kGrow ? 2 : 1;
    const int prev_size = st.prev_size;
    for (int part = 0; part < nparts; part++) {
      RectArgs a;
      bool rect_right;
      a.off_add = off_add; a.rz = clamp16(kZero - st.off); a.corner = 0;
      if (st.dir == kRight) {            // scan_block.rs:147-196
        rect_right = true;
        a.vec_base = si; a.col_base = sj + B - kStep; a.W = kStep; a.H = B;
        a.AD = w.Dc; a.AC = w.Cc; a.OD = w.t1; a.OR_ = w.t2;
        if (st.prev_dir == kDown) a.corner = sat_add(st.D_corner, off_add);
      } else if (st.dir == kDown) {      // scan_block.rs:197-246
        rect_right = false;
        a.vec_base = sj; a.col_base = si + B - kStep; a.W = kStep; a.H = B;
        a.AD = w.Dr; a.AC = w.Rr; a.OD = w.t1; a.OR_ = w.t2;
        if (st.prev_dir == kRight) a.corner = sat_add(st.D_corner, off_add);
      } else if (part == 0) {            // grow, down part: rows si+prev.. x cols sj..sj+prev (scan_block.rs:262-278)
        rect_right = false;
        a.vec_base = sj; a.col_base = si + prev_size; a.W = B - prev_size; a.H = prev_size;
        a.AD = w.Dr; a.AC = w.Rr; a.OD = w.Dc + prev_size; a.OR_ = w.Cc + prev_size;
      } else {                           // grow, right part: rows si..si+B x cols sj+prev..sj+B (scan_block.rs:289-305)
        rect_right = true;
        a.vec_base = si; a.col_base = sj + prev_size; a.W = B - prev_size; a.H = B;
        a.AD = w.Dc; a.AC = w.Cc; a.OD = w.Dr + prev_size; a.OR_ = w.Rr + prev_size;
      }
      const uint32_t vec_len = rect_right ? qlen : rlen, col_len = rect_right ? rlen : qlen;
      a.ncols = a.W;
      if (!XDROP && !m_fqe && a.vec_base + a.H > vec_len) {   // early break (scan_block.rs:1216-1224)
        int lim = (int)col_len - (int)a.col_base;
        if (lim < 0) lim = 0;
        if (lim + 1 < a.ncols) a.ncols = lim + 1;
      }
      a.tw = nullptr; a.tz = nullptr;
      // a shift step finds a new best only if off + max - ZERO > best_max (scan_block.rs:346): below that the argmax of the
      // rectangle is never looked at (the unrelated tails of X-drop alignments: most steps of a block at its maximum size)
      a.key_thr = st.dir == kGrow ? kI16Min - 1 : st.best_max - st.off + kZero;
      a.local = m_local; a.fqs0 = m_fqs && rect_right && a.vec_base == 0;
      a.fqe = m_fqe; a.fq_cls = (int)(qlen % kL); a.fq_row0 = (int)qlen - (int)a.vec_base;
      // First block of an alignment on the packed path. Its input borders are Allocated::clear's MIN = 0 (scan_block.rs:
      // 1322-1339), below the packed path's exact range. Every cell of that block is reachable from the forced origin
      // cell (relative_zero = 16384) through gaps, so each maximum of the recurrence is won by a candidate that descends
      // from the origin; candidates that descend from the borders ("garbage") only ever compare with each other (C of
      // the first column, D' of the first column below row 0) before they lose. Max-plus is shift-equivariant, so
      // starting all borders at S = GL instead of 0 leaves every such comparison, every origin-derived value, every trace
      // bit and the four output borders unchanged -- provided garbage (<= S + W * max score) stays below the smallest
      // origin-derived candidate (>= relative_zero - 3 |open| - (H + W) |extend|), which is tested here. The origin cell
      // itself (scan_block.rs:1130-1132: D00 + score replaced by relative_zero) is the corner operand minus that score.
      bool origin_pk = false;
      if (!PROF && !EXT && P.pk_enable && BA_PK_ORIGIN && st.dir == kGrow && part == 1 && prev_size == 0 && a.vec_base == 0 && a.col_base == 0 && !st.overflow) {
        int GL, GH;
        pk_bounds(a.W, P.gap_open, P.gap_extend, P.pk_smax, GL, GH);
        if (GL + a.W * P.pk_smax < a.rz + 3 * P.gap_open + (a.H + a.W) * P.gap_extend) {
          sc.vec = q; sc.col = r;
          a.off_add = GL; a.corner = a.rz - sc.score(sc.col_token(0), sc.row_token(0));
          origin_pk = true;
        }
      }
      const bool pk_ok = !PROF && !EXT && P.pk_enable && !st.overflow && pk_rect_ok<TALL>(P, a, origin_pk);
      if (origin_pk && !pk_ok) { a.off_add = 0; a.corner = 0; }
      if (TRACE && a.W > 0 && a.H >= 0) {
        const uint32_t woff = st.widx;
        if (m_local && sm.zwords) a.tz = sm.zwords + woff;
        a.tw = trace_push(P, st, sm, rect_right ? a.vec_base : a.col_base, rect_right ? a.col_base : a.vec_base, a.W, a.H, rect_right, lane == 0, pk_ok ? (a.H > 256 ? 2 : 3) : (m_fqe ? 1 : 0));
      }
      add_cells(st, (uint32_t)(a.W * a.H));
      sc.vec = rect_right ? q : r; sc.col = rect_right ? r : q;
      int pbv = 0; unsigned pkey = 15u << kKeyClsShift;
      if (!st.overflow) {
        const bool done = pk_ok;
        if (pk_ok) place_rect_pk<(PROF ? kAA : SCORING), XDROP, TRACE, TALL>(w.smem0, P, P.kc, sc.vec, sc.col, a, w.fr, pbv, pkey, w.ecarry);
#ifdef BA_EMU
        if (wp::lane_id() == 0) { if (done) emu_stats::pk_cells += (uint64_t)a.W * a.H; else emu_stats::exact_cells += (uint64_t)a.W * a.H; }
#endif
        if (done) {}
        else if (PROF && !rect_right) place_rect<SCORING, false, TRACE, XDROP, EXT>(sc, pa, a, w, pbv, pkey);
        else place_rect<SCORING, true, TRACE, XDROP, EXT>(sc, pa, a, w, pbv, pkey);
      }
      if (st.dir == kGrow && part == 0) { gbv = pbv; gbkey = pkey; }
      else { bv = pbv; bkey = pkey; }
    }
    if (st.dir != kGrow) {
      int16_t *o1 = st.dir == kRight ? w.Dr : w.Dc, *o2 = st.dir == kRight ? w.Rr : w.Cc;
      st.D_corner = shift_and_offset(B, o1, o2, w.t1, w.t2, off_add);
    }
    border_maxes(w.Dc, w.Dr, right_max, down_max);
    if (st.dir == kGrow) {
      copy4(B, w.kDc, w.kCc, w.kDr, w.kRr, w.Dc, w.Cc, w.Dr, w.Rr, 0);   // scan_block.rs:315-322
      if (TRACE) { st.ck_widx = st.widx; st.ck_ridx = st.ridx; }
    }
    if (st.overflow) return kRunDone;

    const int this_dir = st.dir;
    st.prev_dir = st.dir;
    const int D_max_max = wp::red_max(bv);
    const int grow_max = wp::red_max(gbv);
    mx = wp::imax(D_max_max, grow_max);
    st.off_max = st.off + mx - kZero;
    st.y_drop_iter++;
    bool grow_no_max = (this_dir == kGrow);

    if (P.step_log && lane == 0) {
      const uint32_t n = *P.step_log_n;
      if (n < P.step_log_cap) {
        StepLog s; s.dir = this_dir; s.i = si; s.j = sj; s.block_size = (uint32_t)B; s.off = st.off;
        s.max = (int16_t)mx; s.right_max = (int16_t)right_max; s.down_max = (int16_t)down_max;
        P.step_log[n] = s;
      }
      *P.step_log_n = n + 1;
    }

    if (st.off_max > st.best_max) {
      if (m_fqe) {   // scan_block.rs:354-368: always in row |q|; the block only grows (first block) or shifts right
        st.best_i = qlen;
        st.best_j = sj + (uint32_t)(this_dir == kRight ? B - kStep : st.prev_size) + (uint32_t)bkey;
      }
      if (XDROP) {
        // decode the argmax (scan_block.rs:370-404)
        const bool use_right = (this_dir != kGrow) || (D_max_max >= grow_max);
        const int m = use_right ? D_max_max : grow_max;
        const int tv = use_right ? bv : gbv;
        const unsigned tk = use_right ? bkey : gbkey;
        const unsigned key = wp::red_max_u(tv == m ? tk : 0u);
        const unsigned cp1 = (key >> kKeyColShift) & 0x3fffu;
        uint32_t v = key & kKeyRowMask, c = 0;
        if (cp1 == 0) v = 0; else c = cp1 - 1;
        if (this_dir == kRight) { st.best_i = si + v; st.best_j = sj + (uint32_t)(B - kStep) + c; }
        else if (this_dir == kDown) { st.best_i = si + (uint32_t)(B - kStep) + c; st.best_j = sj + v; }
        else if (use_right) { st.best_i = si + v; st.best_j = sj + (uint32_t)st.prev_size + c; }
        else { st.best_i = si + (uint32_t)st.prev_size + c; st.best_j = sj + v; }
      }
      if (B < max_size) {
        st.i_ckpt = si; st.j_ckpt = sj; st.off_ckpt = st.off;
        copy4(B, w.kDc, w.kCc, w.kDr, w.kRr, w.Dc, w.Cc, w.Dr, w.Rr, 0);
        if (TRACE) { st.ck_widx = st.widx; st.ck_ridx = st.ridx; }
        grow_no_max = false;
      }
      st.best_max = st.off_max;
      st.y_drop_iter = 0;
    }

    if (XDROP) {
      if (st.off_max < st.best_max - P.x_drop) {
        if (st.x_drop_iter < kXDropIter - 1) st.x_drop_iter++;
        else return kRunDone;
      } else {
        st.x_drop_iter = 0;
      }
    }

    if (si + B > qlen && sj + B > rlen) return kRunDone;
    if (sj + B > rlen) { st.si += kStep; st.dir = kDown; continue; }
    if (si + B > qlen) { st.sj += kStep; st.dir = kRight; continue; }

    if (B * 2 <= max_size) {
      if (st.y_drop_iter > (B / kStep) - 1 || grow_no_max) {
        apply_grow(st, w, TRACE);
        continue;
      }
    }
    }   // !resume_shrink

    if (B > min_size && st.y_drop_iter == 0) {
      const int smx = shrink_max(w.Dc, w.Dr, B);
      if (smx >= mx) {
        st.prev_dir = kGrow;
        const int nb = B / 2;
        st.B = nb;
        copy4(nb, w.Dc, w.Cc, w.Dr, w.Rr, w.Dc, w.Cc, w.Dr, w.Rr, nb);
        st.si += (uint32_t)nb; st.sj += (uint32_t)nb;
        st.i_ckpt = st.si; st.j_ckpt = st.sj; st.off_ckpt = st.off;
        copy4(nb, w.kDc, w.kCc, w.kDr, w.kRr, w.Dc, w.Cc, w.Dr, w.Rr, 0);
        border_maxes(w.Dc, w.Dr, right_max, down_max);
        if (TRACE) { st.ck_widx = st.widx; st.ck_ridx = st.ridx; }
        st.y_drop_iter = 0;
      }
    }

    if (down_max > right_max) { st.si += kStep; st.dir = kDown; }
    else { st.sj += kStep; st.dir = kRight; }
  }
}

// result (scan_block.rs:567-592) + traceback. Borders must be in shared memory.
template <int SCORING, int FLAGS>
BA_DEV void finish_alignment(const Params& P, const AlnState& st, const WarpMem& w, uint32_t slot) {
  constexpr bool TRACE = (FLAGS & kTrace) != 0, XDROP = (FLAGS & kXDrop) != 0;
  const int lane = wp::lane_id();
  wp::syncwarp();
  DevResult res;
  res.status = st.overflow ? (uint32_t)kTraceOverflow : (uint32_t)kOk;
  res.cells = ((uint64_t)st.cells_hi << 32) | st.cells_lo; res.steps = st.steps; res.cigar_n = 0; res.cigar_off = 0;
  if (XDROP || (P.ext_flags & kFreeQueryEndGaps)) {   // scan_block.rs:567-573
    res.score = st.best_max; res.query_idx = st.best_i; res.reference_idx = st.best_j;
  } else if (st.overflow) {
    // aborted mid-way (trace arena full): the block is not at the end of the sequences, there is no result yet
    res.score = 0; res.query_idx = 0; res.reference_idx = 0;
  } else {
    int sv;
    if (st.dir == kRight || st.dir == kGrow) sv = (int)w.Dc[st.qlen - st.si];
    else sv = (int)w.Dr[st.rlen - st.sj];
    res.score = st.off + sv - kZero; res.query_idx = st.qlen; res.reference_idx = st.rlen;
  }
  res.rect_n = st.ridx; res.warp = TRACE ? P.arena_of[st.pair] : slot;
  // TRACE: the trace stays in the pair's arena; ba_traceback_batch_kernel walks it (one walk per lane) after this kernel
  if (lane == 0) {
    P.out[st.pair] = res;
    if (TRACE && st.overflow && P.overflow_list) P.overflow_list[wp::atomic_add(P.overflow_n, 1u)] = st.pair;
  }
  wp::syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Fast phase: several alignments per warp while their block sits at the minimum size (pk_fast_step below). One
// iteration = one Right/Down shift step (scan_block.rs:147-246) of every group of lanes, followed by the post-step
// decisions (scan_block.rs:332-558) as straight-line predicated code. Anything else (grow, early break, values
// outside the packed path's exact range, end of the alignment) parks the group: its status changes and the generic
// phase services it with the whole warp.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Packed fast phase: 32 >> LGT alignments per warp, G = 1 << LGT lanes each (block size 8 * G == min_size),
// borders in registers in the packed layout of ba_packed.cuh (entry e of a border: lane (e mod 4G) / 4,
// register e mod 4, halfword e / 4G). The rectangle of a step is computed with
// pk_cols8; a group whose values leave the packed path's exact range is parked for the generic phase.
// ---------------------------------------------------------------------------------------------
// BIG: the block may be above the minimum size (big_loop): the checkpoint goes straight to the slot's global checkpoint
// in the plain layout, and the shrink test of scan_block.rs:505-510 is evaluated (a shrink parks the group).
template <int SCORING, int FLAGS, int LGT, bool BIG = false>
BA_DEV void pk_fast_step(const Params& P, const WarpMem& w, AlnState& st, PkFast& f, int& status,
                         const uint8_t* qp, const uint8_t* rp, uint32_t my_slot) {
  // The slot's trace arena and rectangle stack are recomputed from the slot index where they are needed: carried
  // across the loop they are 11 more live registers in kernels that already spill (TRACE, 128-register cap).
  SlotMem sm;
  sm.kDc = sm.kCc = sm.kDr = sm.kRr = nullptr; sm.words = nullptr; sm.words_cap = 0; sm.zwords = nullptr; sm.rects = nullptr; sm.rects_cap = 0;
  if ((FLAGS & kTrace) != 0 && status == kStFast) {
    const uint32_t arena = P.arena_of[st.pair];
    sm.words = P.trace_words + (size_t)arena * P.trace_words_per_warp;
    sm.words_cap = P.trace_words_per_warp;
    sm.rects = P.rects + (size_t)arena * P.rects_per_warp;
    sm.rects_cap = P.rects_per_warp;
  }
  constexpr bool XDROP = (FLAGS & kXDrop) != 0, TRACE = (FLAGS & kTrace) != 0;
  constexpr int KIND = (SCORING == kProfile) ? kAA : SCORING;
  constexpr bool PROF = SCORING == kProfile;
  constexpr int G = 1 << LGT;
  constexpr int B = 8 * G;
  const int lane = wp::lane_id(), lg = lane & (G - 1), grp = lane >> LGT;
  const bool active = status == kStFast;
  const bool right = st.dir == kRight;
  const uint32_t si = st.si, sj = st.sj;
  const uint8_t* vec = right ? qp : rp;
  const uint8_t* col = right ? rp : qp;
  const uint32_t vec_base = right ? si : sj;
  const uint32_t col_base = (right ? sj : si) + (B - kStep);

  // ---- step prologue (scan_block.rs:148-158) ----
  const int off = st.off_max;
  const int off_add = clamp16(st.off - off);
  const int corner = (st.prev_dir == (right ? kDown : kRight)) ? sat_add(st.D_corner, off_add) : 0;
  const uint32_t oa2 = pk2(off_add);

  PkScorer<KIND> sc;
  PkpOps po;
  uint2 cw = make_uint2(0u, 0u);
  if (!PROF) {
    sc.init(w.smem0, P);
    sc.rows(*(const uint32_t*)(vec + vec_base + 4 * lg), *(const uint32_t*)(vec + vec_base + 4 * G + 4 * lg));
    cw = BA_CW_PREFETCH ? f.cw : *(const uint2*)(col + col_base);
  } else {
    // sequence-to-profile: qp is the query, the profile plays the reference (rows of a down step, columns of a right step)
    const ProfileDev* pd = P.profiles + st.pair;
    const int8_t* tp = pd->tp;
    po.right = right; po.tlen = pd->tlen; po.hi_off = 4u * G;
    po.gp = pd->gp + 4 * (size_t)col_base;
    po.rres[0] = 0; po.rres[1] = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { po.Bc[k] = 0; po.oR[k] = 0; po.cR[k] = 0; }
    if (right) {
      po.tpc = tp + col_base;
      po.rres[0] = *(const uint32_t*)(qp + vec_base + 4 * lg);
      po.rres[1] = *(const uint32_t*)(qp + vec_base + 4 * G + 4 * lg);
    } else {
      po.tpc = tp + vec_base + 4 * lg;
      cw = BA_CW_PREFETCH ? f.cw : *(const uint2*)(qp + col_base);
      const int ge = P.gap_extend;
      const uint32_t r0 = vec_base + 4 * lg, r1 = r0 + 4 * G;
#pragma unroll
      for (int k = 0; k < 4; k++) {   // get_gap_open_down_C / _R, get_gap_close_down (scan_block.rs:671-676)
        po.Bc[k] = (uint32_t)(((int)wp::ldg(pd->gap_open_R + r0 + k) + ge) + 65536 * ((int)wp::ldg(pd->gap_open_R + r1 + k) + ge));
        po.oR[k] = (uint32_t)((int)wp::ldg(pd->gap_open_C + r0 + k) + 65536 * (int)wp::ldg(pd->gap_open_C + r1 + k));
        po.cR[k] = wp::h_pack((int)wp::ldg(pd->gap_close_C + r0 + k), (int)wp::ldg(pd->gap_close_C + r1 + k));
      }
    }
  }
  // (Tried: touching the next 128-byte line of both sequences every step so that the token loads never leave L1 --
  // 89 % of the kernel's long-scoreboard samples sit at the first use of cw. Measured on B200 it costs 10 % on C2
  // (1156 -> 1042 GCUPS) and 16 % on C3: the extra loads hurt more than the misses they hide. BA_SEQ_PREFETCH re-enables.)
#ifdef BA_SEQ_PREFETCH
  if (lg == 0) wp::touch(vec + vec_base + B + 128);
  if (lg == 1) wp::touch(col + col_base + kStep + 128);
#endif

  constexpr int MCN = PROF ? 4 : PkMc<TRACE, 4>::kN;
  uint32_t D[4], C[4], m[4], mc[MCN] = {};
#pragma unroll
  for (int k = 0; k < 4; k++) { D[k] = wp::vadd2(f.aD[k], oa2); C[k] = wp::vadd2(f.aC[k], oa2); m[k] = 0u; }
  uint32_t* fr = w.fr + grp * 8;
  uint32_t* tw = nullptr;
  if (TRACE && active) tw = trace_push_local(st, sm, right ? si : si + (B - kStep), right ? sj + (B - kStep) : sj, kStep, B, right, lg == 0, 3);
  if constexpr (PROF) pkp_cols8<XDROP, LGT, MCN>(po, P.kc, lg, cw.x, cw.y, D, C, (uint32_t)corner & 0xffffu, m, mc, fr, lg == G - 1);
  else pk_cols8<KIND, XDROP, LGT, TRACE, 4, G>(sc, P.kc, G, lg, cw.x, cw.y, D, C, (uint32_t)corner & 0xffffu, 0, m, mc, fr, lg == G - 1, tw, TRACE && active);
  wp::syncwarp();

  // ---- borders after the step ----
  // D_corner = old orthogonal[STEP-1] + off_add (scan_block.rs:1041-1042): entry 7 = lane 1, register 3, low half
  const int d_corner = sat_add(wp::h_lo((uint32_t)wp::shfl_idx_w((int)f.oD[3], 1, G)), off_add);
  const bool last2 = lg >= G - 2;
  const uint4 fv = *(const uint4*)(fr + (last2 ? 4 * (lg - (G - 2)) : 0));
  const uint32_t fvv[4] = {fv.x, fv.y, fv.z, fv.w};
  const uint32_t selD = last2 ? 0x7632u : 0x3210u, selR = last2 ? 0x5432u : 0x3210u;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    f.aD[k] = D[k]; f.aC[k] = C[k];
    // slide by 8 entries = two lanes; the last two lanes take their high halves from the fresh bottom row
    const uint32_t nd = wp::vadd2((uint32_t)wp::shfl_idx_w((int)f.oD[k], lg + 2, G), oa2);
    const uint32_t nr = wp::vadd2((uint32_t)wp::shfl_idx_w((int)f.oR[k], lg + 2, G), oa2);
    f.oD[k] = wp::prmt(nd, fvv[k], selD);
    f.oR[k] = wp::prmt(nr, fvv[k], selR);
  }
  wp::syncwarp();

  // ---- reductions (scan_block.rs:332-345) ----
  uint32_t lv2 = XDROP ? wp::vmax2(wp::vmax3_2(m[0], m[1], m[2]), m[3]) : m[0];
  int mxv = wp::imax(wp::h_lo(lv2), wp::h_hi(lv2));
#pragma unroll
  for (int s = 0; s < LGT; s++) mxv = wp::imax(mxv, wp::shfl_xor_w(mxv, 1 << s, G));
  // prefix_max over entries 0..7 of both borders (scan_block.rs:1020-1022): low halves of lanes 0 and 1
  const uint32_t a_loc = wp::vmax2(wp::vmax2(f.aD[0], f.aD[1]), wp::vmax2(f.aD[2], f.aD[3]));
  const uint32_t o_loc = wp::vmax2(wp::vmax2(f.oD[0], f.oD[1]), wp::vmax2(f.oD[2], f.oD[3]));
  const uint32_t pm = wp::prmt(a_loc, o_loc, 0x5410u);
  const uint32_t pmm = wp::vmax2((uint32_t)wp::shfl_idx_w((int)pm, 0, G), (uint32_t)wp::shfl_idx_w((int)pm, 1, G));
  const int a_max = wp::h_lo(pmm), o_max = wp::h_hi(pmm);
  const int right_max = right ? a_max : o_max, down_max = right ? o_max : a_max;
  // suffix_max over the last two entries of both borders (scan_block.rs:1030-1032): high halves of registers 2 and 3 of the last lane
  int smx = 0;
  if (BIG) {
    const uint32_t sl = wp::vmax2(wp::vmax2(f.aD[2], f.aD[3]), wp::vmax2(f.oD[2], f.oD[3]));
    smx = wp::h_hi((uint32_t)wp::shfl_idx_w((int)sl, G - 1, G));
  }
  unsigned key = 0;
  if (XDROP) {
    key = pk_lane_key<4, MCN>(m, mc, lg, G, mxv);
#pragma unroll
    for (int s = 0; s < LGT; s++) { const unsigned u = (unsigned)wp::shfl_xor_w((int)key, 1 << s, G); key = u > key ? u : key; }
  }

  // (no step log here: a batch that asks for it runs without a fast phase, ba_runtime.cu)

  // ---- post-step decisions (scan_block.rs:332-558 restricted to B == min_size, shift steps) ----
  if (active) {
#ifdef BA_EMU
    if (lg == 0) { if (BIG) emu_stats::big_cells += (uint64_t)kStep * B; else emu_stats::fast_steps++; }
#endif
    st.off = off;
    st.steps++;
    add_cells(st, (uint32_t)(kStep * B));
    st.prev_dir = st.dir;
    st.D_corner = d_corner;
    st.off_max = off + mxv - kZero;
    st.y_drop_iter++;
    if (st.off_max > st.best_max) {
      if (XDROP) {
        const unsigned cp1 = (key >> kKeyColShift) & 0x3fffu;
        uint32_t v = key & kKeyRowMask, c = 0;
        if (cp1 == 0) v = 0; else c = cp1 - 1;
        if (right) { st.best_i = si + v; st.best_j = sj + (B - kStep) + c; }
        else { st.best_i = si + (B - kStep) + c; st.best_j = sj + v; }
      }
      if (B < (int)P.max_size) {
        st.i_ckpt = si; st.j_ckpt = sj; st.off_ckpt = off;
        // checkpoint copy of all four borders (scan_block.rs:413-420): kept in shared memory in the packed
        // layout, array order D_col, C_col, D_row, R_row; moved to the slot's global checkpoint when the group parks
        if (BIG) {
          int16_t *ad = right ? w.kDc : w.kDr, *ac = right ? w.kCc : w.kRr, *od = right ? w.kDr : w.kDc, *orr = right ? w.kRr : w.kCc;
          pk_store4(ad, lg, G, f.aD); pk_store4(ac, lg, G, f.aC); pk_store4(od, lg, G, f.oD); pk_store4(orr, lg, G, f.oR);
        } else {
          uint4* ck = (uint4*)w.ck + (grp * 4) * G + lg;
          const int ia = right ? 0 : 2 * G, io = right ? 2 * G : 0;
          ck[ia] = make_uint4(f.aD[0], f.aD[1], f.aD[2], f.aD[3]); ck[ia + G] = make_uint4(f.aC[0], f.aC[1], f.aC[2], f.aC[3]);
          ck[io] = make_uint4(f.oD[0], f.oD[1], f.oD[2], f.oD[3]); ck[io + G] = make_uint4(f.oR[0], f.oR[1], f.oR[2], f.oR[3]);
          st.ck_pk = 1u;
        }
        if (TRACE) { st.ck_widx = st.widx; st.ck_ridx = st.ridx; }
      }
      st.best_max = st.off_max;
      st.y_drop_iter = 0;
    }
    int nstatus = kStFast;
    if (XDROP) {
      if (st.off_max < st.best_max - P.x_drop) {
        if (st.x_drop_iter < kXDropIter - 1) st.x_drop_iter++;
        else nstatus = kStDone;
      } else {
        st.x_drop_iter = 0;
      }
    }
    if (TRACE && st.overflow) nstatus = kStDone;
    if (nstatus == kStFast) {
      if (si + B > st.qlen && sj + B > st.rlen) nstatus = kStDone;
      else if (sj + B > st.rlen) { st.si = si + kStep; st.dir = kDown; }
      else if (si + B > st.qlen) { st.sj = sj + kStep; st.dir = kRight; }
      else if (2 * B <= (int)P.max_size && st.y_drop_iter > (B / kStep) - 1) nstatus = kStNeedGrow;   // state left untouched
      else if (BIG && B > (int)P.min_size && st.y_drop_iter == 0 && smx >= mxv) nstatus = kStNeedShrink;   // scan_block.rs:505-510, done by run_generic
      else if (down_max > right_max) { st.si = si + kStep; st.dir = kDown; }
      else { st.sj = sj + kStep; st.dir = kRight; }
      if (nstatus == kStFast && !shift_eligible<SCORING, XDROP>(st, B)) nstatus = kStNeedGeneric;
      if (TRACE && nstatus == kStFast && !trace_room(st, sm.words_cap, sm.rects_cap, B)) nstatus = kStNeedGeneric;
    }
    status = nstatus;
  }
  // ---- range guard for the next packed step (ba_packed.cuh): all borders + next off_add inside [GL, GH] ----
  {
    int GL, GH;
    pk_bounds(kStep, P.gap_open, P.gap_extend, P.pk_smax, GL, GH);
    if (PROF) { GL = kPkpGL; GH = kPkpGH; }
    const int noa = clamp16(st.off - st.off_max);
    const int lo_b = wp::imax(GL - noa, kI16Min), hi_b = wp::imin(GH - noa, kI16Max);
    // Only the D registers need the test: C <= D and R <= D bound the gap tables from above, and from below they
    // only have to stay clear of i16 wrap-around, which holds by induction: every C / R value is >= 0 when it is
    // produced (or checked by pk_borders_ok on entry) and then moves by at most |noa| <= 1024 per step for the
    // at most G <= 8 steps it stays in a border.
    bool ok = lo_b <= hi_b && noa >= -1024 && noa <= 1024 && pk_in_range_d(f.aD, f.oD, lo_b, hi_b);
    if (st.prev_dir != kGrow && st.prev_dir != st.dir) {
      const int cv = sat_add(st.D_corner, noa);
      ok = ok && cv >= 0 && cv <= GH;
    }
    const unsigned badm = wp::ballot(!ok);
    if (status == kStFast && ((badm >> (lane & ~(G - 1))) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u))) != 0u) status = kStNeedGeneric;
  }
  // the registers are laid out for the executed direction; if the next fast step goes the other way the
  // borders swap roles
  {
    const uint32_t flip = (status == kStFast && st.dir != st.prev_dir) ? 0xffffffffu : 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) {   // one bit-select (LOP3) per register
      const uint32_t aD = f.aD[k], oD = f.oD[k], aC = f.aC[k], oR = f.oR[k];
      f.aD[k] = (aD & ~flip) | (oD & flip); f.oD[k] = (oD & ~flip) | (aD & flip);
      f.aC[k] = (aC & ~flip) | (oR & flip); f.oR[k] = (oR & ~flip) | (aC & flip);
    }
  }
  if (BA_CW_PREFETCH && status == kStFast) pk_fast_prefetch<SCORING, LGT>(f, st, qp, rp);
}

BA_DEV void bcast_state(AlnState& d, const AlnState& s, int src) {
#define BA_BC(F) d.F = (decltype(d.F))wp::shfl_idx((int)s.F, src)
  BA_BC(pair); BA_BC(qlen); BA_BC(rlen); BA_BC(si); BA_BC(sj); BA_BC(B); BA_BC(prev_size); BA_BC(dir); BA_BC(prev_dir);
  BA_BC(off); BA_BC(off_max); BA_BC(best_max); BA_BC(best_i); BA_BC(best_j); BA_BC(i_ckpt); BA_BC(j_ckpt);
  BA_BC(off_ckpt); BA_BC(y_drop_iter); BA_BC(x_drop_iter); BA_BC(D_corner); BA_BC(cells_lo); BA_BC(cells_hi); BA_BC(steps);
  BA_BC(widx); BA_BC(ridx); BA_BC(ck_widx); BA_BC(ck_ridx); BA_BC(overflow); BA_BC(ck_pk);
#undef BA_BC
}

// ---------------------------------------------------------------------------------------------
// Warp main: carve shared memory, then keep up to four alignments in flight until the ticket
// counter runs dry.
// ---------------------------------------------------------------------------------------------
// gb: the four live borders of the generic phase are in global memory (P.gborders) instead of shared memory. At
// max block sizes >= 1024 they are 8-32 KB per warp and would cap the SM at 8 warps or fewer; the generic phase
// touches only the first B entries, and the fast phase keeps its borders in registers anyway.
BA_HD size_t warp_smem_bytes(uint32_t max_size, bool gb = false) {
  const size_t ms = max_size < 32 ? 32 : max_size;
  size_t b = (gb ? 0 : 4 * ms * sizeof(int16_t)) + 2 * 16 * sizeof(int16_t) + 4 * sizeof(int32_t) + 64 * sizeof(uint32_t) + 512 * sizeof(uint32_t);
  if (max_size > 32) b += ms;   // ecarry (rectangles swept in several chunks, TRACE)
  return (b + 15) & ~(size_t)15;
}

BA_DEV void bind_slot(const Params& P, uint32_t slot, WarpMem& w, SlotMem& sm, bool trace) {
  const size_t ms = P.max_size < 32 ? 32 : P.max_size;
  int16_t* g = P.ckpt + (size_t)slot * 4 * ms;
  sm.kDc = g; sm.kCc = g + ms; sm.kDr = g + 2 * ms; sm.kRr = g + 3 * ms;
  w.kDc = sm.kDc; w.kCc = sm.kCc; w.kDr = sm.kDr; w.kRr = sm.kRr;
  sm.words = nullptr; sm.words_cap = 0; sm.zwords = nullptr; sm.rects = nullptr; sm.rects_cap = 0;
  (void)trace;
}
// the trace arena of the pair being serviced
BA_DEV void bind_trace(const Params& P, uint32_t pair, SlotMem& sm) {
  const uint32_t arena = P.arena_of[pair];
  sm.words = P.trace_words + (size_t)arena * P.trace_words_per_warp;
  sm.words_cap = P.trace_words_per_warp;
  sm.zwords = P.trace_zwords ? P.trace_zwords + (size_t)arena * P.trace_words_per_warp : nullptr;
  sm.rects = P.rects + (size_t)arena * P.rects_per_warp;
  sm.rects_cap = P.rects_per_warp;
}

// FM selects the fast phase: 0 = none (every step in the generic phase), 16 + LGT = packed fast phase with
// 1 << LGT lanes per alignment (min block size 8 << LGT); + 16 (FM >= 32): live borders in global memory.
template <int SCORING, int FLAGS, int FM>
BA_DEV void warp_main(const Params& P, unsigned char* smem, int warp_in_block, uint32_t warp_global) {
  constexpr bool TRACE = (FLAGS & kTrace) != 0;
  constexpr bool GB = FM >= 32;
  constexpr bool PKF = FM >= 16;
  constexpr int LGT = PKF ? (FM & 15) : 5;     // log2(lanes per alignment group)
  constexpr int GW = 1 << LGT;
  const int lane = wp::lane_id();
  const size_t ms = P.max_size < 32 ? 32 : P.max_size;
  WarpMem w;
  w.mat = (const int8_t*)smem;                     // header: matrix, then the packed-path tables
  w.smem0 = smem;
  unsigned char* base = smem + kSmemHeader + (size_t)warp_in_block * warp_smem_bytes(P.max_size, GB);
  int16_t* p16 = (int16_t*)base;
  if (GB) {
    int16_t* g16 = P.gborders + (size_t)warp_global * 4 * ms;
    w.Dc = g16; w.Cc = g16 + ms; w.Dr = g16 + 2 * ms; w.Rr = g16 + 3 * ms;
  } else {
    w.Dc = p16; p16 += ms; w.Cc = p16; p16 += ms; w.Dr = p16; p16 += ms; w.Rr = p16; p16 += ms;
  }
  w.t1 = p16; p16 += 16; w.t2 = p16; p16 += 16;
  w.misc = (int32_t*)p16; p16 += 8;
  w.fr = (uint32_t*)p16; p16 += 128;
  w.ck = (uint32_t*)p16; p16 += 1024;
  w.ecarry = (uint8_t*)p16;
  w.kDc = w.kCc = w.kDr = w.kRr = nullptr;
  const uint32_t spw = P.slots_per_warp;

  // Groups of GW lanes; every lane carries the state of its group's alignment. Without a fast phase
  // (profiles, extended modes, min block size the fast phases do not serve) only group 0 is used and every
  // alignment runs start to end in the generic phase.
  const int my_g = lane >> LGT;
  AlnState st;
  st.pair = 0; st.qlen = 0; st.rlen = 0; st.si = 0; st.sj = 0; st.B = 32; st.prev_size = 0; st.dir = kRight; st.prev_dir = kGrow;
  st.off = 0; st.off_max = 0; st.best_max = 0; st.best_i = 0; st.best_j = 0; st.i_ckpt = 0; st.j_ckpt = 0; st.off_ckpt = 0;
  st.y_drop_iter = 0; st.x_drop_iter = 0; st.D_corner = 0; st.cells_lo = 0; st.cells_hi = 0; st.steps = 0;
  st.widx = 0; st.ridx = 0; st.ck_widx = 0; st.ck_ridx = 0; st.overflow = 0; st.ck_pk = 0;
  PkFast pf;
#pragma unroll
  for (int k = 0; k < 4; k++) { pf.aD[k] = 0; pf.aC[k] = 0; pf.oD[k] = 0; pf.oR[k] = 0; }
  pf.cw = make_uint2(0u, 0u);
  int status = kStEmpty;
  const uint8_t* qp = P.seq;
  const uint8_t* rp = P.seq;
  const uint32_t my_slot = warp_global * spw + ((uint32_t)my_g < spw ? (uint32_t)my_g : 0u);
  bool tickets_left = true;

  for (;;) {
    // ---- service: refill empty groups, run the generic phase for parked ones ----
    for (int g = 0; g < (int)spw; g++) {
      const int sg = wp::shfl_idx(status, g * GW);
      if (sg == kStFast) continue;
      if (sg == kStEmpty && !tickets_left) continue;
      SlotMem sm;
      const uint32_t slot = warp_global * spw + (uint32_t)g;
      bind_slot(P, slot, w, sm, TRACE);
      AlnState gs;
      const bool mine = my_g == g;
      if (sg == kStEmpty) {
        uint32_t t = 0;
        if (lane == 0) t = wp::atomic_add(P.ticket, 1u);
        t = (uint32_t)wp::shfl_idx((int)t, 0);
        if (t >= P.n_pairs) { tickets_left = false; continue; }
        init_alignment<SCORING, FLAGS>(P, gs, P.order ? P.order[t] : t, w);
        if (TRACE) bind_trace(P, gs.pair, sm);
      } else {
        bcast_state(gs, st, g * GW);
        if (TRACE) bind_trace(P, gs.pair, sm);
        // a parked group's registers are still laid out for the step it executed last (= prev_dir)
        if (PKF) {
          pk_fast_spill<LGT>(pf, w, gs.prev_dir, mine);
          if (gs.ck_pk) { pk_ckpt_flush<LGT>(w, sm, g, mine); gs.ck_pk = 0u; }
        }
        if (sg == kStNeedGrow) apply_grow(gs, w, TRACE);
      }
      int r = kRunDone;
      if (sg != kStDone) r = run_generic<SCORING, FLAGS, (FM >= 32)>(P, gs, w, sm, slot);
      if (r == kRunDone) {
        finish_alignment<SCORING, FLAGS>(P, gs, w, slot);
        if (mine) status = kStEmpty;
        g--;            // try to refill this slot right away
        continue;
      }
      wp::syncwarp();
      if (PKF) pk_fast_load<LGT>(pf, w, gs.dir, mine);
      if (mine) {
        st = gs; status = kStFast;
        qp = P.seq + P.q_off[gs.pair];
        if (SCORING != kProfile) rp = P.seq + P.r_off[gs.pair];
        if (PKF && BA_CW_PREFETCH) pk_fast_prefetch<SCORING, LGT>(pf, st, qp, rp);
      }
      wp::syncwarp();
    }
    if (wp::ballot(status == kStFast) == 0u) break;
    // ---- fast phase: run until some group needs the generic phase ----
    for (;;) {
      if (PKF) pk_fast_step<SCORING, FLAGS, LGT>(P, w, st, pf, status, qp, rp, my_slot);
      else break;
      if (wp::ballot(status != kStFast && status != kStEmpty) != 0u) break;
      if (wp::ballot(status == kStFast) == 0u) break;
    }
  }
}

// Traceback from an arbitrary end position of a pair that already ran (legacy block_cigar_* calls):
// the trace of pair `pair` is still in the arena of the slot that aligned it.
BA_DEV void warp_traceback(const Params& P, const uint8_t* lut, uint32_t pair, uint32_t qi, uint32_t rj, bool eq, DevResult* out1) {
  DevResult res = P.out[pair];
  const uint32_t slot = P.arena_of[pair];
  const uint32_t* words = P.trace_words + (size_t)slot * P.trace_words_per_warp;
  const Rect* rects = P.rects + (size_t)slot * P.rects_per_warp;
  uint32_t* runs = P.run_scratch + (size_t)slot * P.runs_per_warp;
  const uint8_t* q = P.seq + P.q_off[pair];
  const uint8_t* r = P.profiles ? nullptr : P.seq + P.r_off[pair];
  res.status = (uint32_t)kOk;
  const uint32_t* zwords = ((P.ext_flags & kLocalStart) && P.trace_zwords) ? P.trace_zwords + (size_t)slot * P.trace_words_per_warp : nullptr;
  emit_cigar(P, lut, words, zwords, (P.ext_flags & kFreeQueryStartGaps) != 0, rects, res.rect_n, qi, rj, q, r, eq, runs, res);
  if (wp::lane_id() == 0) *out1 = res;
}

// Batch traceback: lane `t` of the launch walks the trace of pair list[t] from the end position of its result
// (Trace::cigar / cigar_eq, scan_block.rs:1469-1480): 32 independent dependent-load chains per warp instead of one.
// The runs come out reversed (Cigar::add order, cigar.rs:71-79) into the arena's scratch; the warp then copies each
// lane's list forward into the output stream (space claimed with one atomicAdd per alignment).
// Prefetch distance in rectangle records, into L2: with 128 walks per SM the windows of an L1 prefetch (8 sectors per
// rectangle) do not fit the L1 (measured: no gain); in L2 (126 MB) the windows of every walk of the batch do.
constexpr uint32_t kTbLaneAhead = 8;
// STAGED: eight lanes per walk (t is valid in the first lane of each group of eight; launch_traceback_batch). The group
// copies the trace words of the rectangle the walk will enter next into shared memory (`stage`: two buffers of
// kTbStageWords words per group) with asynchronous copies while the walk is still inside the current rectangle, so
// that a cell step reads its word from shared memory instead of following a chain of L2 / DRAM round trips (one per
// column of every rectangle: 1200 cycles per cell step on C5, ncu r02: long_scoreboard 4.9 stalls per issue). Only
// packed-path rectangles of at most kTbStageWords words are staged (shift steps of blocks 32 .. 128); anything larger is
// read from global memory as before.
constexpr uint32_t kTbStageWords = 128;
BA_DEV bool tb_stageable(const Rect& rc, const uint32_t* pw) {
  return ((rc.right >> 1) & 3u) == 3u && (uint32_t)(rc.h >> 3) * rc.w <= kTbStageWords && (((uintptr_t)pw) & 15u) == 0u;
}
template <bool STAGED>
BA_DEV void lanes_traceback(const Params& P, const uint8_t* lut, const uint32_t* list, uint32_t n, uint32_t t, uint32_t* stage = nullptr) {
  const int lane = wp::lane_id();
  const bool have = t < n;
  uint32_t pair = 0, nruns = 0, ok = 0;
  DevResult res;
  res.score = 0; res.query_idx = 0; res.reference_idx = 0; res.status = (uint32_t)kNotRun; res.cells = 0; res.steps = 0;
  res.cigar_n = 0; res.cigar_off = 0; res.rect_n = 0; res.warp = 0;
  const uint32_t* runs = nullptr;
  bool active_walk = false, fqs_ = false;
  const uint32_t *words_ = nullptr, *zwords_ = nullptr;
  const Rect* rects_ = nullptr;
  uint32_t* rs_ = nullptr;
  const uint8_t *q_ = nullptr, *r_ = nullptr;
  WalkState s;
  s.i = 0; s.j = 0; s.ridx = 0; s.table = 0; s.cur_op = 0; s.cur_len = 0; s.nruns = 0; s.bad = 0; s.stop = 0;
  if (have) {
    pair = list ? list[t] : t;
    res = P.out[pair];
    res.cigar_n = 0; res.cigar_off = 0;
    if (res.status == (uint32_t)kOk) {
      const uint32_t arena = P.arena_of[pair];
      const uint32_t* words = P.trace_words + (size_t)arena * P.trace_words_per_warp;
      const uint32_t* zwords = ((P.ext_flags & kLocalStart) && P.trace_zwords) ? P.trace_zwords + (size_t)arena * P.trace_words_per_warp : nullptr;
      const Rect* rects = P.rects + (size_t)arena * P.rects_per_warp;
      uint32_t* rs = P.run_scratch + (size_t)arena * P.runs_per_warp;
      const uint8_t* q = P.seq + P.q_off[pair];
      const uint8_t* r = P.profiles ? nullptr : P.seq + P.r_off[pair];
      const bool fqs = (P.ext_flags & kFreeQueryStartGaps) != 0;
      active_walk = true;
      words_ = words; zwords_ = zwords; rects_ = rects; rs_ = rs; q_ = q; r_ = r; fqs_ = fqs;
      s.i = res.query_idx; s.j = res.reference_idx; s.ridx = res.rect_n;
    }
  }
  // Flat loop, one iteration = one record fetch or one cell step of every lane: nested per-lane loops would let the 32
  // walks drift apart and serialise (measured: 116 ms for 10 k walks of ~90 k steps, i.e. ~2400 cycles per step).
  {
    Rect rc;
    rc.row = 0xffffffffu; rc.col = 0xffffffffu; rc.h = 0; rc.w = 0; rc.right = 0; rc.word_off = 0;   // "no rectangle yet"
    const uint32_t* tw = nullptr;
    const bool eq = P.cigar_eq != 0;
    const uint32_t cap = P.runs_per_warp;
    // STAGED state: stg = the words of rc are in stage buffer `cur`; pend = index of the record whose words are being
    // copied into the other buffer (0xffffffff: none), n_rec = that record
    bool stg = false;
    uint32_t cur = 0, pend = 0xffffffffu;
    Rect n_rec = rc, nn_rec = rc;
    uint32_t nn_idx = 0xffffffffu;
    const int hl = lane & 7;
    uint32_t* sbuf = STAGED ? stage + (size_t)(lane >> 3) * 2 * kTbStageWords : nullptr;
    for (;;) {
      const bool go = active_walk && (s.i > 0 || s.j > 0) && !s.bad && !s.stop;
      if (wp::ballot(go) == 0u) break;
      const bool in_rect = s.i >= rc.row && s.j >= rc.col;
      if (STAGED) {
        // ---- a walk needs its next record: the whole group of eight lanes takes part ----
        const bool need = go && !in_rect;
        const bool gneed = wp::shfl_idx_w((int)need, 0, 8) != 0;
        if (wp::ballot(gneed) != 0u) {
          {   // every lane of the warp runs this block (its shuffles use the full mask); groups without `need` copy nothing
            // walker: the newest rectangle containing (i, j) (scan_block.rs:1578-1590), one record per iteration
            uint32_t use_pend = 0, do_stage = 0, nw = 0;
            const uint32_t* pw = nullptr;
            if (need) {
              if (s.ridx == 0) s.bad = 1u;
              else {
                s.ridx--;
                use_pend = pend == s.ridx ? 1u : 0u;
                rc = use_pend ? n_rec : rects_[s.ridx];
                tw = rect_words_ptr(words_, P.trace_pool, rc);
                if (s.i >= rc.row && s.j >= rc.col && tb_stageable(rc, tw)) { do_stage = 1u; pw = tw; nw = (uint32_t)(rc.h >> 3) * rc.w; }
                stg = false;
              }
            }
            do_stage = (uint32_t)wp::shfl_idx_w((int)do_stage, 0, 8);
            use_pend = (uint32_t)wp::shfl_idx_w((int)use_pend, 0, 8);
            // the group's copy in flight (if any) has landed: its buffer may be read or refilled. (Only the group with the
            // event waits: the other groups' copies were issued moments ago.)
            if (gneed) wp::cp_async_wait_all();
            {
              // not predicted (first rectangle of the walk, or records were skipped): copy now. (The shuffles are executed by
              // every lane of the warp; only the copies depend on the group's flags.)
              const uint64_t pa = (uint64_t)(uintptr_t)pw;
              const uint32_t plo = (uint32_t)wp::shfl_idx_w((int)(uint32_t)pa, 0, 8), phi = (uint32_t)wp::shfl_idx_w((int)(uint32_t)(pa >> 32), 0, 8);
              const uint32_t* src = (const uint32_t*)(uintptr_t)(((uint64_t)phi << 32) | plo);
              nw = (uint32_t)wp::shfl_idx_w((int)nw, 0, 8);
              const uint32_t gcur = (uint32_t)wp::shfl_idx_w((int)cur, 0, 8);
              if (do_stage && !use_pend) {
                for (uint32_t c4 = (uint32_t)hl; 4 * c4 < nw; c4 += 8) wp::cp_async16(sbuf + (gcur ^ 1u) * kTbStageWords + 4 * c4, src + 4 * c4);
                wp::cp_async_commit();
                wp::cp_async_wait_all();
              }
            }
            // next record down the stack: its words go into the buffer the walk has just left. The record itself was
            // loaded one event ago (nn_rec), the one after it is requested now: no load of this block is waited for here
            uint32_t nstage = 0, nnw = 0;
            const uint32_t* npw = nullptr;
            if (need && !s.bad) {
              if (do_stage) { cur ^= 1u; stg = true; }
              pend = 0xffffffffu;
              if (s.ridx > 0) {
                n_rec = (nn_idx == s.ridx - 1) ? nn_rec : rects_[s.ridx - 1];
                npw = rect_words_ptr(words_, P.trace_pool, n_rec);
                if (tb_stageable(n_rec, npw)) { nstage = 1u; nnw = (uint32_t)(n_rec.h >> 3) * n_rec.w; pend = s.ridx - 1; }
                nn_idx = 0xffffffffu;
                if (s.ridx > 1) { nn_rec = rects_[s.ridx - 2]; nn_idx = s.ridx - 2; }
                // further down: bring trace words and records from DRAM into L2. The stack is walked in order and shift
                // rectangles lie back to back in the arena, so the words kTbLaneAhead rectangles down are (about) that many
                // rectangle sizes below this one -- an address computed without loading that record
                const uint32_t rw = (uint32_t)(rc.h >> 3) * rc.w;
                if (!(rc.right & kRectPool) && rw <= kTbStageWords && rc.word_off >= (kTbLaneAhead + 1) * rw) {
                  const uint32_t* ppw = tw - kTbLaneAhead * rw;
#pragma unroll
                  for (uint32_t u = 0; u < kTbStageWords; u += 8u) if (u < rw) wp::touch_l2(ppw + u);
                }
                if (s.ridx >= 4 * kTbLaneAhead && (s.ridx & 1u) == 0u) wp::touch_l2(rects_ + (s.ridx - 4 * kTbLaneAhead));
              }
            }
            nstage = (uint32_t)wp::shfl_idx_w((int)nstage, 0, 8);
            {
              const uint64_t pa = (uint64_t)(uintptr_t)npw;
              const uint32_t plo = (uint32_t)wp::shfl_idx_w((int)(uint32_t)pa, 0, 8), phi = (uint32_t)wp::shfl_idx_w((int)(uint32_t)(pa >> 32), 0, 8);
              const uint32_t* src = (const uint32_t*)(uintptr_t)(((uint64_t)phi << 32) | plo);
              nnw = (uint32_t)wp::shfl_idx_w((int)nnw, 0, 8);
              const uint32_t gcur = (uint32_t)wp::shfl_idx_w((int)cur, 0, 8);
              if (nstage) {
                for (uint32_t c4 = (uint32_t)hl; 4 * c4 < nnw; c4 += 8) wp::cp_async16(sbuf + (gcur ^ 1u) * kTbStageWords + 4 * c4, src + 4 * c4);
                wp::cp_async_commit();
              }
            }
          }
          wp::syncwarp();
        }
      }
      if (go && (!STAGED || in_rect)) {
        if (!in_rect) {
          // the newest rectangle containing (i, j) (scan_block.rs:1578-1590): one record per iteration
          if (s.ridx == 0) s.bad = 1u;
          else {
            s.ridx--;
            rc = rects_[s.ridx];
            tw = rect_words_ptr(words_, P.trace_pool, rc);
            // the walk is a chain of dependent loads over words written long ago: fetch the words of the rectangle
            // kTbLaneAhead records down the stack (the walk visits the stack in order) and, further ahead, the records
            if (s.ridx >= kTbLaneAhead) {
              const Rect pr = rects_[s.ridx - kTbLaneAhead];
              const uint32_t* pw = rect_words_ptr(words_, P.trace_pool, pr);
              // a missing load brings in one 32-byte sector: touch every sector of a shift rectangle (one per column at
              // block size 64), the first eight of anything larger
              const uint32_t nw = (uint32_t)(pr.h >> 3) * pr.w;
#pragma unroll
              for (uint32_t u = 0; u < 64u; u += 8u) if (u < nw) wp::touch_l2(pw + u);
              if (s.ridx >= 4 * kTbLaneAhead && (s.ridx & 1u) == 0u) wp::touch_l2(rects_ + (s.ridx - 4 * kTbLaneAhead));
            }
          }
        } else if (fqs_ && (rc.right & 1u) && s.i == 0) {
          s.stop = 1u;      // FREE_QUERY_START_GAPS: stop on row 0, which always lies in right rectangles (scan_block.rs:1597-1600)
        } else {
          const bool rc_right = (rc.right & 1u) != 0;
          const uint32_t layout = (rc.right >> 1) & 3u;
          const uint32_t v = rc_right ? s.i - rc.row : s.j - rc.col;
          const uint32_t c = rc_right ? s.j - rc.col : s.i - rc.row;
          uint32_t nib;
          bool zstop = false;
          if (layout >= 2u) {
            const uint32_t G = (uint32_t)rc.h >> 3;
            const uint32_t chn = layout == 2u ? (v >> 8) : 0u, v8 = layout == 2u ? (v & 255u) : v, hh = layout == 2u ? 128u : ((uint32_t)rc.h >> 1);
            const uint32_t half = v8 >= hh ? 1u : 0u, vv = v8 - half * hh;
            const uint32_t wi = c * G + chn * 32u + (vv >> 2);
#ifdef BA_EMU
            if (STAGED && stg) emu_stats::tb_staged++; else emu_stats::tb_global++;
#endif
            const uint32_t word = (STAGED && stg) ? sbuf[cur * kTbStageWords + wi] : tw[wi];
            nib = (word >> (16u * half + 4u * (vv & 3u))) & 15u;
          } else {
            const uint32_t R = layout == 1u ? 8u : (uint32_t)rect_rows_per_lane(rc.h);
            const uint32_t CH = 32u * R, ngroups = (uint32_t)rc.w >> 3;
            const uint32_t ch = v / CH, ln = (v % CH) / R, k = v % R;
            const size_t widx = (((size_t)ch * ngroups + (c >> 3)) * R + k) * 32 + ln;
            // LOCAL_START: the alignment starts at a cell equal to relative_zero (scan_block.rs:1606-1612)
            if (zwords_ && s.table == 0 && ((zwords_[rc.word_off + widx] >> (c & 7)) & 1u)) zstop = true;
            nib = (tw[widx] >> (4 * (c & 7))) & 15u;
          }
          if (zstop) s.stop = 1u;
          else {
            const uint32_t e = lut[(rc_right ? 64u : 0u) | (s.table << 4) | nib];
            uint32_t op = e & 7u;
            const uint32_t di = (e >> 3) & 1u, dj = (e >> 4) & 1u;
            if (op == 1u && eq) op = (q_[s.i] == r_[s.j]) ? 2u : 3u;   // scan_block.rs:1620-1628
            if ((di && s.i == 0) || (dj && s.j == 0)) s.bad = 1u;
            else {
              s.i -= di; s.j -= dj; s.table = e >> 5;
              if (op == s.cur_op) s.cur_len++;
              else {
                if (s.cur_len) { if (s.nruns < cap) rs_[s.nruns] = (s.cur_len << 4) | s.cur_op; s.nruns++; }
                s.cur_op = op; s.cur_len = 1;
              }
            }
          }
        }
      }
    }
  }
  if (active_walk) {
    if (s.cur_len) { if (s.nruns < P.runs_per_warp) rs_[s.nruns] = (s.cur_len << 4) | s.cur_op; s.nruns++; }
    nruns = s.nruns; runs = rs_;
    ok = (!s.bad && s.nruns <= P.runs_per_warp) ? 1u : 0u;
    if (!ok) res.status = (uint32_t)kCigarOverflow;
  }
  wp::syncwarp();
  for (int l = 0; l < 32; l++) {
    if (!wp::shfl_idx((int)ok, l)) continue;
    const uint32_t nl = (uint32_t)wp::shfl_idx((int)nruns, l);
    const uint64_t rp = (uint64_t)(uintptr_t)runs;
    const uint32_t plo = (uint32_t)wp::shfl_idx((int)(uint32_t)(rp & 0xffffffffu), l), phi = (uint32_t)wp::shfl_idx((int)(uint32_t)(rp >> 32), l);
    const uint32_t* src = (const uint32_t*)(uintptr_t)(((uint64_t)phi << 32) | plo);
    unsigned long long base = 0;
    if (lane == 0) base = wp::atomic_add64(P.cigar_used, (unsigned long long)nl);
    const uint32_t blo = (uint32_t)wp::shfl_idx((int)(uint32_t)(base & 0xffffffffu), 0), bhi = (uint32_t)wp::shfl_idx((int)(uint32_t)(base >> 32), 0);
    base = ((unsigned long long)bhi << 32) | blo;
    const bool fits = base + nl <= P.cigar_cap;
    if (fits) for (uint32_t u = (uint32_t)lane; u < nl; u += 32) P.cigar_stream[base + u] = src[nl - 1 - u];
    if (lane == l) {
      if (fits) { res.cigar_n = nl; res.cigar_off = base; }
      else res.status = (uint32_t)kCigarOverflow;
    }
  }
  if (have) P.out[pair] = res;
}

}  // namespace ba
