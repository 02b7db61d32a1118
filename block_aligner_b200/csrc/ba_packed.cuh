// ba_packed.cuh -- the DP of one rectangle on packed 2 x i16 DPX instructions.
//
// Included by ba_kernel.cuh. Same recurrence as place_rect_r (reference: place_block, scan_block.rs:1083-1228;
// prefix scan avx2.rs:297-338), but two rows per 32-bit register: a group of G lanes (G = H / 8, H = 32..256 rows;
// taller rectangles: a stack of 256-row chunks, place_rect_pk_tall) owns one rectangle; lane lg keeps rows 4*lg + k in the low halfword and rows 4*G + 4*lg + k in the high
// halfword of register k (k = 0..3). Every add/max of the recurrence is one VIADDMNMX.S16x2 / VIMNMX.S16x2,
// i.e. two cells per ALU instruction.
//
// Exactness. The reference computes with *saturating* i16 adds (_mm256_adds_epi16); the packed DPX adds wrap.
// The two agree whenever no intermediate leaves the i16 range, and the reference's scan "phantoms" (negative
// constants, avx2.rs:321-337) and its T[top-1] = 0 carry-in cannot win whenever every scanned value is >= 0.
// Both are guaranteed by a range guard on the rectangle's inputs (pk_bounds): with every input D/C value in
// [GL, GH], GL = -(W*open + open - extend) and GH = 32767 - W*max(matrix), every cell of a W-column rectangle
// obeys 0 <= D + (open - extend) and D <= 32767 (D falls by at most |open| and rises by at most max(matrix) per
// column). A rectangle whose inputs fail the guard is computed by the exact 32-bit path (place_rect_r) instead;
// in practice that is the first block of an alignment (borders start at MIN = 0) and little else.
//
// The vertical-gap scan is kept in "U space", U = T - (open - extend): U[r] = max(U[r-1] + ext, D'[r]) needs no
// add for x = D' + (open - extend), and D = max(D', U + (open - extend)) is one add-max.
#pragma once

namespace ba {

BA_DEV uint32_t pk2(int v) { return wp::h_pack(v, v); }

#ifndef BA_PK_UNROLL
#define BA_PK_UNROLL 1
#endif
// Tuning switches (tools/build_variant.sh; measurements in profiles/r02_variants.txt).
// BA_PK_IMAD = 1: the plain adds of the column loop are issued as IMAD (x * 1 + y, the 1 opaque to the compiler). OFF:
// measured neutral (ptxas already issues them as VIADD, which runs on the FMA pipe like IMAD).
#ifndef BA_PK_IMAD
#define BA_PK_IMAD 0
#endif
#define BA_PK_ONE 1u
// BA_PK_SPLIT_MC = 1 (default): the X-drop column trackers of the two half-blocks live in separate registers
// (mc[k], mc[K + k]) and are updated with SEL; 0: both in one register, updated with PRMT. Same pipe, but the split
// form needs no (column + 1) operand in PRMT position: measured +1.5 % on B200 / C2 (1195 -> 1213 GCUPS, r02). TRACE
// kernels (which spill at the 128-register cap) and the profile loop keep the four packed trackers.
#ifndef BA_PK_SPLIT_MC
#define BA_PK_SPLIT_MC 1
#endif
// BA_PK_KVAR: rows-per-lane variants of the generic phase's packed rectangles (bit 0: one register per lane for 32 / 64
// rows, bit 1: two registers per lane for 128 rows; 0: four registers per lane everywhere, fewer lanes busy). Each
// variant is one more copy of the rectangle code, and the kernels live at the edge of the SM's 32 KB instruction cache:
// first measured as a loss (fewer instructions, sm__icc_request_hit_rate 96 -> 87 % on C2, 78 -> 68 % on C3), a gain
// (C2 +2 %, C3 +5 %) once the rest of the generic phase had been made 12 KB smaller. BA_PK_GT = 1: rectangles that fill the
// warp (128 / 256 rows with both variants on) use a compile-time group size of 32 (column loop 111 -> 98 instructions);
// BA_PK_GT = 2 adds such a copy for 64-row rectangles (loop 70 -> 57) and is a loss again (C3 -8 %).
#ifndef BA_PK_KVAR
#define BA_PK_KVAR 3
#endif
#ifndef BA_PK_GT
#define BA_PK_GT 1
#endif
// BA_PK_TALL = 1: rectangles taller than 256 rows on the packed path, in chained 256-row chunks (place_rect_pk_tall), in
// the kernels that have a fast phase and max block >= 1024. Measured on C5 (block 64..=2048, TRACE | X_DROP): kernel
// 345 -> 373 GCUPS, instruction-cache hit rate 91 -> 98 % (profiles/r02_tall_ab.txt, r02_ncu_k_C5_metrics.csv).
#ifndef BA_PK_TALL
#define BA_PK_TALL 1
#endif
template <bool TRACE, int K> struct PkMc { static constexpr bool kSplit = BA_PK_SPLIT_MC && !TRACE; static constexpr int kN = kSplit ? 2 * K : K; };
constexpr int kPkUnroll = BA_PK_UNROLL;    // 1, 2 or 4

// shared-memory scoring tables of the packed path (one copy per CTA, built by stage_tables)
constexpr int kMatBytes = 1024;            // raw matrix (exact path)
constexpr int kPkTabBytes = 8320;          // kNuc: packed score pairs, bank-spread layout (nuc_tab_index); kAA: [27][32] i16
constexpr int kTbLutBytes = 128;           // traceback step table (tb_entry, ba_kernel.cuh)
constexpr int kSmemHeader = kMatBytes + kPkTabBytes + kTbLutBytes;

// Step table: the reference's 64-entry OP_LUT (scan_block.rs:1508-1572) folded per rectangle orientation.
// index = nibble (bit 0 D == A, bit 1 D == B, bit 2 A opened here, bit 3 B opened at the row above; A = the gap
// table along the rectangle's sequential columns, B = the one along its vectors) | table << 4 | right << 6;
// value = op | di << 3 | dj << 4 | next table << 5   (tables: 0 = D, 1 = C, 2 = R; ops: 1 M, 4 I, 5 D)
BA_HD uint8_t tb_entry(uint32_t idx) {
  const uint32_t nib = idx & 15u, table = (idx >> 4) & 3u, right = (idx >> 6) & 1u;
  const uint32_t t = nib & 3u, t2 = nib >> 2;
  const uint32_t tabA = right ? 1u : 2u, tabB = right ? 2u : 1u, opA = right ? 5u : 4u, opB = right ? 4u : 5u;
  uint32_t op, ntab;
  if (table == tabA) { op = opA; ntab = (t2 & 1u) ? 0u : tabA; }
  else if (table == tabB) { op = opB; ntab = (t2 & 2u) ? 0u : tabB; }
  else if (t == 0) { op = 1u; ntab = 0u; }
  else if (t & 1u) { op = opA; ntab = (t2 & 1u) ? 0u : tabA; }      // C wins over R when both equal D (scan_block.rs:1539-1542)
  else { op = opB; ntab = (t2 & 2u) ? 0u : tabB; }
  const uint32_t di = (op == 5u) ? 0u : 1u, dj = (op == 4u) ? 0u : 1u;
  return (uint8_t)(op | (di << 3) | (dj << 4) | (ntab << 5));
}

template <int KIND> struct PkScorer;
// NucMatrix (scores.rs:195-209): row (c & 7) * 16, column b & 15. One table word holds the scores of a pair of rows
// (tokens lo, hi) against one column class. Layout (nuc_tab_index): word = class * 260 + r(lo, hi), where the low four
// bits of r are bits 1-2 of the two row tokens -- the bits that tell A, C, G, T apart ('A' & 15 = 1, 'C' = 3, 'G' = 7,
// 'T' = 4). With the plain [class][lo][hi] order the 32 lanes of a column step hit 8 of the 32 shared-memory banks
// (6.1 wavefronts per LDS on random ACGT, ncu: 6.5 G bank conflicts per C2 launch); with this order 2.1.
constexpr int kNucClsStride = 260;         // words per column class (256 + 4: neighbouring classes start 4 banks apart)
BA_HD uint32_t nuc_row_index(uint32_t lo, uint32_t hi) {
  return ((lo >> 1) & 3u) | (((hi >> 1) & 3u) << 2) | ((lo & 1u) << 4) | (((lo >> 3) & 1u) << 5) | ((hi & 1u) << 6) | (((hi >> 3) & 1u) << 7);
}
BA_HD uint32_t nuc_tab_index(uint32_t cls, uint32_t lo, uint32_t hi) { return cls * kNucClsStride + nuc_row_index(lo, hi); }
template <> struct PkScorer<kNuc> {
  const unsigned char* tab; uint32_t rt[4];
  BA_DEV void init(const unsigned char* smem, const Params&) { tab = smem + kMatBytes; }
  BA_DEV void rows(uint32_t wlo, uint32_t whi) {
    // nuc_row_index of the four (lo, hi) token pairs at once, one per byte
    const uint32_t a = (wlo >> 1) & 0x03030303u, b = (whi << 1) & 0x0c0c0c0cu;
    const uint32_t c = (wlo << 4) & 0x10101010u, d = (wlo << 2) & 0x20202020u;
    const uint32_t e = (whi << 6) & 0x40404040u, f = (whi << 4) & 0x80808080u;
    const uint32_t t = a | b | c | d | e | f;
#pragma unroll
    for (int k = 0; k < 4; k++) rt[k] = ((t >> (8 * k)) & 0xffu) << 2;
  }
  BA_DEV uint32_t colh(uint32_t cb) const { return cb & 7u; }
  // class * stride + row offset: one IMAD (FMA pipe; the ALU pipe is the column loop's bottleneck)
  BA_DEV uint32_t score(uint32_t ch, int k) const { return *(const uint32_t*)(tab + (ch * (uint32_t)(4 * kNucClsStride) + rt[k])); }
};
// AAMatrix (scores.rs:110-127): row c * 32, column b & 31
template <> struct PkScorer<kAA> {
  const unsigned char* tab; uint32_t rt[4];
  BA_DEV void init(const unsigned char* smem, const Params&) { tab = smem + kMatBytes; }
  BA_DEV void rows(uint32_t wlo, uint32_t whi) {
#pragma unroll
    for (int k = 0; k < 4; k++) rt[k] = (((wlo >> (8 * k)) & 31u) << 1) | ((((whi >> (8 * k)) & 31u) << 1) << 16);
  }
  BA_DEV uint32_t colh(uint32_t cb) const { return cb << 6; }
  BA_DEV uint32_t score(uint32_t ch, int k) const {
    const uint32_t lo = *(const uint16_t*)(tab + ch + (rt[k] & 0xffffu));
    const uint32_t hi = *(const uint16_t*)(tab + ch + (rt[k] >> 16));
    return lo | (hi << 16);
  }
};
// ByteMatrix (scores.rs:263-267)
template <> struct PkScorer<kByte> {
  uint32_t rt[4]; uint32_t match, mismatch;
  BA_DEV void init(const unsigned char*, const Params& P) { match = (uint32_t)(int)P.matrix[0] & 0xffffu; mismatch = (uint32_t)(int)P.matrix[1] & 0xffffu; }
  BA_DEV void rows(uint32_t wlo, uint32_t whi) {
#pragma unroll
    for (int k = 0; k < 4; k++) rt[k] = ((wlo >> (8 * k)) & 0xffu) | (((whi >> (8 * k)) & 0xffu) << 16);
  }
  BA_DEV uint32_t colh(uint32_t cb) const { return cb; }
  BA_DEV uint32_t score(uint32_t ch, int k) const {
    const uint32_t lo = ((rt[k] & 0xffffu) == ch) ? match : mismatch;
    const uint32_t hi = ((rt[k] >> 16) == ch) ? match : mismatch;
    return lo | (hi << 16);
  }
};

// Build the packed-path tables from the staged matrix. tid / nthreads: the calling thread's share.
template <int SCORING>
BA_DEV void stage_tables(unsigned char* smem, int tid, int nthreads) {
  const int8_t* mat = (const int8_t*)smem;
  for (int i = tid; i < kTbLutBytes; i += nthreads) smem[kMatBytes + kPkTabBytes + i] = tb_entry((uint32_t)i);
  if (SCORING == kNuc) {
    uint32_t* t = (uint32_t*)(smem + kMatBytes);
    for (int i = tid; i < 2048; i += nthreads) {
      const int cls = i >> 8, a = (i >> 4) & 15, b = i & 15;
      t[nuc_tab_index((uint32_t)cls, (uint32_t)a, (uint32_t)b)] = wp::h_pack((int)mat[cls * 16 + a], (int)mat[cls * 16 + b]);
    }
  } else if (SCORING == kAA) {
    int16_t* t = (int16_t*)(smem + kMatBytes);
    for (int i = tid; i < 27 * 32; i += nthreads) t[i] = (int16_t)mat[i];
  }
}

// input range [GL, GH] under which a W-column packed rectangle is exact (see the header comment)
BA_DEV void pk_bounds(int W, int go, int ge, int smax, int& GL, int& GH) {
  GL = -(W * go + (go - ge));
  GH = kI16Max - W * smax;
}

// Eight consecutive columns of a packed rectangle. D/C: the previous column on entry, the last column on exit.
// m / mc: per-row running maximum and the (column + 1) of its last occurrence (X-drop argmax, scan_block.rs:
// 1194-1201); without XDROP only m[0] is kept (block maximum). fr[c] receives the bottom row of column c
// (row 8G - 1) as T | D << 16 from the group's last lane.
//
// TRACE: one word per lane and column (Rect layout 3, word index = column * G + lane in group): nibble k of the low
// 16 bits belongs to row 4lg + k, nibble k of the high 16 bits to row 4G + 4lg + k; bit 0 D == C, bit 1 D == R,
// bit 2 C opened in this cell, bit 3 the R gap of this row was opened at the row above (the reference's trace /
// trace2 words, scan_block.rs:1166-1190). Every comparison is an equality between x <= y, computed without
// predicates as a halfword mask (pk_eqmask).
// per halfword 0xffff where y == x, else 0; needs x <= y (then y + ~x = y - x - 1 is -1 exactly for equal halves)
BA_DEV uint32_t pk_eqmask(uint32_t y, uint32_t x) { return wp::viaddmin2(y, ~x, 0u); }

// K = registers per lane and array (4, 2 or 1): a rectangle of H rows is spread over G = H / (2 K) lanes, lane lg keeps
// rows K lg + k in the low halfwords and rows K G + K lg + k in the high halfwords (k < K). K = 4 is the layout of the
// fast phase and of 256-row rectangles; K = 2 / 1 put a 128 / 64-row rectangle on all 32 lanes (32 rows: 16 lanes)
// instead of 16 / 8 / 4 of them -- the grow rectangles and the shift steps of blocks 64 and 128 (C3: 42 % of the
// kernel's instructions ran in that loop at 1/4 of the lanes, ncu r02). NST = Kogge-Stone stages executed (>= log2 G;
// a stage whose shuffle distance reaches G returns the lane's own value and changes nothing because extend < 0).
// GT: the group size when it is known at compile time (0: run-time `Gr`) -- shuffle widths become immediates instead of
// operands that the generic phase, short of registers, rebuilds in every iteration of the column loop.
// CHAIN: the rectangle continues one above it (rectangles taller than 256 rows are swept in 256-row chunks, chunk-major):
// `tops[c]` = R | D << 16 of the row just above for the eight columns (the bottom row the chunk above wrote to `fr`),
// `dprev` = that row's D one column to the left of the first (the diagonal input of column 0), `etop` / `eout` = the
// "R gap opened at this row" bit of that row and of this chunk's last row (TRACE), `tstride` = trace words per column.
template <int KIND, bool XDROP, int NST, bool TRACE = false, int K = 4, int GT = 0, bool CHAIN = false>
BA_DEV void pk_cols8(const PkScorer<KIND>& sc, const PkConst& kc, int Gr, int lg, uint32_t cw0, uint32_t cw1,
                     uint32_t (&D)[K], uint32_t (&C)[K], uint32_t corner_lo, int cbase, uint32_t (&m)[K], uint32_t (&mc)[PkMc<TRACE, K>::kN],
                     uint32_t* fr, bool writer, uint32_t* tw = nullptr, bool tstore = false, int ncol8 = 8,
                     const uint32_t* tops = nullptr, uint32_t dprev = 0u, const uint8_t* etop = nullptr, uint8_t* eout = nullptr,
                     int tstride = 0) {
  // ncol8 < 8: global-mode early break inside this group of eight columns (scan_block.rs:1216-1224): only the first
  // ncol8 columns exist; D / C then hold the last computed column, like the reference's D_col / C_col after its break
  static_assert(!TRACE || K == 4, "trace words are laid out for four registers per lane");
  const int G = GT ? GT : Gr;
  constexpr bool SPLIT = PkMc<TRACE, K>::kSplit;
  // The uniform constants are copied into vector registers once per call (opaque_zero): ptxas otherwise rebuilds
  // every packed constant from its 16-bit halves at each use.
  const uint32_t z = wp::opaque_zero();
  const uint32_t ge2 = kc.ge2 + z, or2 = kc.or2 + z;
  const uint32_t kge1 = kc.kge[1] + z, kge2 = kc.kge[2] + z, kge3 = kc.kge[3] + z;
  const uint32_t lanedec = (uint32_t)(lg * K) * kc.ge1 + (lg ? 0x10000u : 0u);   // packed K * lg * gap_extend
  const uint32_t one = z + BA_PK_ONE;
  // kPkUnroll columns per iteration of the rolled loop: the kernel's hot code has to stay inside the SM's
  // instruction cache (ncu: sm__icc_request_hit_rate), which a fully unrolled 8-column body does not
  uint64_t cwq = ((uint64_t)cw1 << 32) | cw0;
#pragma unroll 1
  for (int h = 0; h < ncol8 / kPkUnroll; h++) {
    const uint32_t cwh = (uint32_t)cwq;
    cwq >>= 8 * kPkUnroll;
#pragma unroll
    for (int cc = 0; cc < kPkUnroll; cc++) {
      const int cidx = h * kPkUnroll + cc;
      const uint32_t ch = sc.colh((cwh >> (8 * cc)) & 0xffu);
      // diagonal input of the lane's first rows: the previous column of the row above. Lane 0: low half = the
      // rectangle's corner (first column only; MIN = 0 afterwards, scan_block.rs:1211), high half = row K G - 1,
      // which is the low half of the last lane's last register.
      uint32_t up = (uint32_t)wp::shfl_idx_w((int)D[K - 1], lg - 1, G);
      uint32_t topw = 0u;
      if (CHAIN) {
        topw = tops[cidx];
        if (lg == 0) up = (up << 16) | (dprev & 0xffffu);
        dprev = topw >> 16;                       // the next column's diagonal input
      } else {
        if (lg == 0) up = (up << 16) | ((cbase + cidx == 0) ? corner_lo : 0u);
      }
      uint32_t dd[K], c11[K], uu[K];
      uint32_t acc[K];                            // trace nibbles of the lane's rows (TRACE)
#pragma unroll
      for (int k = 0; k < K; k++) {
        const uint32_t s2 = sc.score(ch, k);
        const uint32_t d00 = k ? D[k - 1] : up;
        // packed add as one 32-bit add: no borrow, both halves >= |open| (guard)
        const uint32_t c11o = BA_PK_IMAD ? D[k] * one + kc.go1 : D[k] + kc.go1;
        c11[k] = wp::viaddmax2(C[k], ge2, c11o);
        if (TRACE) acc[k] = pk_eqmask(c11[k], c11o) & 0x00040004u;       // C opened here <=> C == D10 + open
        dd[k] = wp::viaddmax2(d00, s2, c11[k]);
        uu[k] = k ? wp::viaddmax2(uu[k - 1], ge2, dd[k]) : dd[0];
      }
      // Kogge-Stone over the lane aggregates, both half-blocks at once; stage s decays by (K << s) * extend
      uint32_t inc = uu[K - 1];
      // CHAIN: the U of the row above enters the scan through the first lane's aggregate (low half; the high half's 0 + K extend < 0 cannot win)
      if (CHAIN && lg == 0) inc = wp::viaddmax2((topw - kc.or2) & 0xffffu, kge3, inc);
#pragma unroll
      for (int s = 0; s < NST; s++) {
        const uint32_t u = (uint32_t)wp::shfl_up_w((int)inc, 1 << s, G);
        const int mult = K << s;                  // 1, 2, 3, 4 -> ge2 / kge[]; 4 << t -> dec[t]
        const uint32_t dc = mult == 1 ? ge2 : (mult == 2 ? kge1 : (mult == 4 ? kge3 : kc.dec[s + (K == 4 ? 0 : (K == 2 ? -1 : -2))] + z));
        inc = wp::viaddmax2(u, dc, inc);
      }
      // carry into the lane: the lanes above (same half) and, for the high half, the whole low half-block
      const uint32_t tl = (uint32_t)wp::shfl_idx_w((int)inc, G - 1, G);
      uint32_t ex = (uint32_t)wp::shfl_up_w((int)inc, 1, G);
      // first lane: nothing above (T[top - 1] = 0, which cannot win); CHAIN: the U of the row above = its R - (open - extend)
      if (lg == 0) ex = CHAIN ? ((topw - kc.or2) & 0xffffu) : 0u;
      const uint32_t cin = wp::viaddmax2(tl << 16, lanedec, ex);
      uint32_t UnL = 0;
      uint32_t eprev = 0u;                        // "R gap opened at this row" (T == x) of the previous row of the lane
#pragma unroll
      for (int k = 0; k < K; k++) {
        const uint32_t Un = wp::viaddmax2(cin, k == 0 ? ge2 : (k == 1 ? kge1 : (k == 2 ? kge2 : kge3)), uu[k]);
        uint32_t Dn;
        if (TRACE) {
          const uint32_t Rv = wp::vadd2(Un, or2);
          Dn = wp::vmax2(Rv, dd[k]);
          acc[k] |= pk_eqmask(Dn, c11[k]) & 0x00010001u;                 // D == C
          acc[k] |= pk_eqmask(Dn, Rv) & 0x00020002u;                     // D == R
          if (k > 0) acc[k] |= eprev & 0x00080008u;
          eprev = pk_eqmask(Un, dd[k]);                                  // T == x  <=>  U == D'
        } else {
          Dn = wp::viaddmax2(Un, or2, dd[k]);
        }
        if (k == K - 1) UnL = Un;
        if (XDROP) {
          bool ph, pl;
          m[k] = wp::vibmax2(Dn, m[k], ph, pl);
          // kept in a vector register (opaque_zero) so that PRMT can take its selector as an immediate
          const uint32_t c1 = (uint32_t)(cbase + cidx + 1) + wp::opaque_zero();
          if (SPLIT) {
            mc[k] = pl ? c1 : mc[k];
            mc[PkMc<TRACE, K>::kN - K + k] = ph ? c1 : mc[PkMc<TRACE, K>::kN - K + k];
          } else {
            if (pl) mc[k] = wp::prmt(mc[k], c1, 0x3254u);
            if (ph) mc[k] = wp::prmt(mc[k], c1, 0x5410u);
          }
        } else {
          m[0] = wp::vmax2(m[0], Dn);
        }
        D[k] = Dn; C[k] = c11[k];
      }
      if (TRACE) {
        // bit 3 of the lane's first rows comes from the last row of the lane above: lane 0's high half continues the
        // last lane's low half, lane 0's low half is the top of the rectangle (no row above: R01 = MIN)
        uint32_t eup = (uint32_t)wp::shfl_idx_w((int)eprev, lg - 1, G);
        if (lg == 0) eup = (eup << 16) | ((CHAIN && etop && etop[cidx]) ? 0xffffu : 0u);
        acc[0] |= eup & 0x00080008u;
        const uint32_t word = acc[0] | (acc[1 % K] << 4) | (acc[2 % K] << 8) | (acc[3 % K] << 12);
        if (tstore) tw[(size_t)(cbase + cidx) * (CHAIN ? tstride : G) + lg] = word;
        if (CHAIN && writer) eout[cidx] = (uint8_t)(eprev >> 31);
      }
      if (writer) fr[cidx] = wp::prmt(wp::vadd2(UnL, or2), D[K - 1], 0x7632u);   // T.hi | D.hi << 16
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Sequence-to-profile (place_block_profile_right / _down, scan_block.rs:612-783) on the packed datapath, for the fast
// phase. One body serves both directions; the operands of the direction that does not apply are zero:
//   right (vectors = query rows, columns = profile positions): per column  A = gap_open_C + ext, cl = gap_close_C,
//         oRc = gap_open_R  (one 16-byte load of ProfileDev::gp), scores = 8 words of ProfileDev::tp per four columns
//         (row of the lane's query residue, the four positions), one PRMT per pair of cells;
//   down  (vectors = profile positions, columns = query residues): per row  Bc = gap_open_R + ext, oR = gap_open_C,
//         cR = gap_close_C (the reference swaps the roles, scan_block.rs:671-676), scores = two words of tp per column
//         (row of the column's residue, the lane's four + four positions).
// Recurrence per cell (T space: the open cost varies per row or column, so the U-space trick of pk_cols8 does not apply):
//   C11 = max(C10 + ext, D10 + A + Bc);  D' = max(D00 + s, C11 + cl);  x = D' + oRc + oR;  T = max(T_up + ext, x);
//   D = max(D', T + cR).  Plain 32-bit adds stand in for packed adds wherever a negative operand is added to halves that
// the range guard keeps >= its magnitude (no borrow between the halves); operands of those adds are "v * 65537".
// Exact while every border value lies in [kPkpGL, kPkpGH] (pk_fast_step's guard): per column D falls by at most
// |open| + |ext| + |close| <= 384 and rises by at most score + close_C + close_R <= 381, whatever the profile holds.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPkpGL = 8 * 384 + 128 + 128, kPkpGH = kI16Max - 8 * 381;
struct PkpOps {
  bool right;
  const uint32_t* gp;              // right: gp + 4 * (first column position)
  const int8_t* tpc;               // right: tp + first column position; down: tp + first row position of the lane (low half)
  uint32_t tlen, hi_off;           // row stride of tp; down: byte offset of the lane's high-half rows (4 * G)
  uint32_t rres[2];                // right: query residues of the lane's rows (low / high half), one per byte
  uint32_t Bc[4], oR[4], cR[4];    // down: per-row operands (Bc, oR: v * 65537 form; cR: packed halves); right: 0
};

template <bool XDROP, int LGT, int MC>      // MC >= 4: only mc[0..3] are used (both column trackers packed in one register)
BA_DEV void pkp_cols8(const PkpOps& o, const PkConst& kc, int lg, uint32_t cw0, uint32_t cw1,
                      uint32_t (&D)[4], uint32_t (&C)[4], uint32_t corner_lo, uint32_t (&m)[4], uint32_t (&mc)[MC],
                      uint32_t* fr, bool writer) {
  constexpr int G = 1 << LGT;
  const uint32_t z = wp::opaque_zero();
  const uint32_t ge2 = kc.ge2 + z;
  const uint32_t kge1 = kc.kge[1] + z, kge2 = kc.kge[2] + z, kge3 = kc.kge[3] + z;
  const uint32_t lanedec = (uint32_t)lg * kc.lane1 + (lg ? 0x10000u : 0u);   // packed 4 * lg * gap_extend
  uint64_t cwq = ((uint64_t)cw1 << 32) | cw0;
  uint32_t w[8] = {};
#pragma unroll 1
  for (int cidx = 0; cidx < 8; cidx++) {
    uint32_t A = 0, cl = 0, oRc = 0, wl = 0, wh = 0;
    if (o.right) {
      if ((cidx & 3) == 0) {
        // the lane's eight rows x four columns: row r of tp is the residue of query row r
#pragma unroll
        for (int k = 0; k < 4; k++) {
          w[k] = *(const uint32_t*)(o.tpc + (size_t)((o.rres[0] >> (8 * k)) & 0xffu) * o.tlen + cidx);
          w[4 + k] = *(const uint32_t*)(o.tpc + (size_t)((o.rres[1] >> (8 * k)) & 0xffu) * o.tlen + cidx);
        }
      }
      const uint4 g = *(const uint4*)(o.gp + 4 * cidx);
      A = g.x; cl = g.y; oRc = g.z;
    } else {
      const int8_t* row = o.tpc + (size_t)((uint32_t)cwq & 0xffu) * o.tlen;
      wl = *(const uint32_t*)row;
      wh = *(const uint32_t*)(row + o.hi_off);
    }
    cwq >>= 8;
    // selector of column (cidx & 3): byte c of the low word sign-extended into the low half, of the high word into the high half
    const uint32_t c4 = (uint32_t)(cidx & 3);
    const uint32_t selc = (c4 | ((8u | c4) << 4) | ((4u | c4) << 8) | ((12u | c4) << 12));
    uint32_t up = (uint32_t)wp::shfl_idx_w((int)D[3], lg - 1, G);
    if (lg == 0) up = (up << 16) | (cidx == 0 ? corner_lo : 0u);
    uint32_t dd[4], c11[4], tt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t selk = ((uint32_t)k | ((8u | (uint32_t)k) << 4) | ((4u | (uint32_t)k) << 8) | ((12u | (uint32_t)k) << 12));
      const uint32_t s2 = o.right ? wp::prmt_sx(w[k], w[4 + k], selc) : wp::prmt_sx(wl, wh, selk);
      const uint32_t d00 = k ? D[k - 1] : up;
      const uint32_t c11o = D[k] + A + o.Bc[k];
      c11[k] = wp::viaddmax2(C[k], ge2, c11o);
      dd[k] = wp::viaddmax2(d00, s2, c11[k] + cl);
      const uint32_t x = dd[k] + oRc + o.oR[k];
      tt[k] = k ? wp::viaddmax2(tt[k - 1], ge2, x) : x;
    }
    uint32_t inc = tt[3];
#pragma unroll
    for (int s = 0; s < LGT; s++) {
      const uint32_t u = (uint32_t)wp::shfl_up_w((int)inc, 1 << s, G);
      inc = wp::viaddmax2(u, s == 0 ? kge3 : kc.dec[s] + z, inc);
    }
    const uint32_t tl = (uint32_t)wp::shfl_idx_w((int)inc, G - 1, G);
    uint32_t ex = (uint32_t)wp::shfl_up_w((int)inc, 1, G);
    if (lg == 0) ex = 0u;
    const uint32_t cin = wp::viaddmax2(tl << 16, lanedec, ex);
    uint32_t Tn3 = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t Tn = wp::viaddmax2(cin, k == 0 ? ge2 : (k == 1 ? kge1 : (k == 2 ? kge2 : kge3)), tt[k]);
      const uint32_t Dn = wp::viaddmax2(Tn, o.cR[k], dd[k]);
      if (k == 3) Tn3 = Tn;
      if (XDROP) {
        bool ph, pl;
        m[k] = wp::vibmax2(Dn, m[k], ph, pl);
        const uint32_t c1 = (uint32_t)(cidx + 1) + wp::opaque_zero();
        if (pl) mc[k] = wp::prmt(mc[k], c1, 0x3254u);
        if (ph) mc[k] = wp::prmt(mc[k], c1, 0x5410u);
      } else {
        m[0] = wp::vmax2(m[0], Dn);
      }
      D[k] = Dn; C[k] = c11[k];
    }
    if (writer) fr[cidx] = wp::prmt(Tn3, D[3], 0x7632u);   // T.hi | D.hi << 16
  }
}

// Best cell of the lane under the reference's order: value desc, AVX lane (row mod 16) asc, column desc, row desc
// (scan_block.rs:1194-1201, avx2.rs:271-274). pk_lane_key: key of the lane's best cell among those equal to M
// (0 if none); key format as in place_rect_r: (15 - class) << 28 | (column + 1) << 14 | row. Every row has seen
// a cell >= 0 (guard), so there is no "no cell" case. The two rows of a register share their class (K G is a multiple
// of 16). MC = K: both column trackers packed in mc[k]; MC = 2 K: low halves in mc[k], high halves in mc[K + k].
template <int K, int MC>
BA_DEV unsigned pk_lane_key(const uint32_t (&m)[K], const uint32_t (&mc)[MC], int lg, int G, int M) {
  const uint32_t M2 = pk2(M);
  const unsigned base0 = ((15u - (unsigned)((K * lg) & 15)) << kKeyClsShift) | (unsigned)(K * lg);
  unsigned key = 0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    bool ph, pl;
    wp::vibmax2(m[k], M2, ph, pl);            // m >= M, i.e. m == M for M = the maximum
    const unsigned base = base0 - ((unsigned)k << kKeyClsShift) + (unsigned)k;
    const unsigned clo = MC == K ? (mc[k] & 0xffffu) : mc[k];
    const unsigned chi = MC == K ? (mc[k] >> 16) : mc[(MC == K ? 0 : K) + k];
    const unsigned klo = base | (clo << kKeyColShift);
    const unsigned khi = (base + (unsigned)(K * G)) | (chi << kKeyColShift);
    if (pl && klo > key) key = klo;
    if (ph && khi > key) key = khi;
  }
  return key;
}
template <int K>
BA_DEV int pk_lane_max(const uint32_t (&m)[K]) {
  uint32_t v = m[0];
#pragma unroll
  for (int k = 1; k < K; k++) v = wp::vmax2(v, m[k]);
  return wp::imax(wp::h_lo(v), wp::h_hi(v));
}

// borders in shared memory <-> packed registers of lane lg (rows K lg.. and K G + K lg..)
template <int K>
BA_DEV void pk_loadk(const int16_t* p, int lg, int G, uint32_t (&r)[K]) {
  if (K == 4) {
    const uint2 lo = *(const uint2*)(p + 4 * lg);
    const uint2 hi = *(const uint2*)(p + 4 * G + 4 * lg);
    r[0] = wp::prmt(lo.x, hi.x, 0x5410u); r[1 % K] = wp::prmt(lo.x, hi.x, 0x7632u);
    r[2 % K] = wp::prmt(lo.y, hi.y, 0x5410u); r[3 % K] = wp::prmt(lo.y, hi.y, 0x7632u);
  } else if (K == 2) {
    const uint32_t lo = *(const uint32_t*)(p + 2 * lg), hi = *(const uint32_t*)(p + 2 * G + 2 * lg);
    r[0] = wp::prmt(lo, hi, 0x5410u); r[1 % K] = wp::prmt(lo, hi, 0x7632u);
  } else {
    r[0] = (uint32_t)(uint16_t)p[lg] | ((uint32_t)(uint16_t)p[G + lg] << 16);
  }
}
template <int K>
BA_DEV void pk_storek(int16_t* p, int lg, int G, const uint32_t (&r)[K]) {
  if (K == 4) {
    uint2 lo, hi;
    lo.x = wp::prmt(r[0], r[1 % K], 0x5410u); lo.y = wp::prmt(r[2 % K], r[3 % K], 0x5410u);
    hi.x = wp::prmt(r[0], r[1 % K], 0x7632u); hi.y = wp::prmt(r[2 % K], r[3 % K], 0x7632u);
    *(uint2*)(p + 4 * lg) = lo;
    *(uint2*)(p + 4 * G + 4 * lg) = hi;
  } else if (K == 2) {
    *(uint32_t*)(p + 2 * lg) = wp::prmt(r[0], r[1 % K], 0x5410u);
    *(uint32_t*)(p + 2 * G + 2 * lg) = wp::prmt(r[0], r[1 % K], 0x7632u);
  } else {
    p[lg] = (int16_t)(r[0] & 0xffffu);
    p[G + lg] = (int16_t)(r[0] >> 16);
  }
}
BA_DEV void pk_load4(const int16_t* p, int lg, int G, uint32_t (&r)[4]) { pk_loadk<4>(p, lg, G, r); }
BA_DEV void pk_store4(int16_t* p, int lg, int G, const uint32_t (&r)[4]) { pk_storek<4>(p, lg, G, r); }

// lane-local range test of the D registers of both borders (the packed fast phase's per-step guard)
BA_DEV bool pk_in_range_d(const uint32_t (&a)[4], const uint32_t (&b)[4], int lo, int hi) {
  uint32_t mn = wp::vmin3_2(a[0], a[1], a[2]), mx = wp::vmax3_2(a[0], a[1], a[2]);
  mn = wp::vmin3_2(mn, a[3], b[0]); mx = wp::vmax3_2(mx, a[3], b[0]);
  mn = wp::vmin3_2(mn, b[1], b[2]); mx = wp::vmax3_2(mx, b[1], b[2]);
  mn = wp::vmin2(mn, b[3]); mx = wp::vmax2(mx, b[3]);
  bool p0, p1, p2, p3;
  wp::vibmax2(mn, pk2(lo), p0, p1);
  wp::vibmax2(pk2(hi), mx, p2, p3);
  return p0 && p1 && p2 && p3;
}

// lane-local range test of packed values: every halfword of every register within [lo, hi]
template <int N>
BA_DEV bool pk_in_range(const uint32_t (&a)[N], const uint32_t (&b)[N], int lo, int hi) {
  uint32_t mn = a[0], mx = a[0];
#pragma unroll
  for (int k = 1; k < N; k++) { mn = wp::vmin2(mn, a[k]); mx = wp::vmax2(mx, a[k]); }
#pragma unroll
  for (int k = 0; k < N; k++) { mn = wp::vmin2(mn, b[k]); mx = wp::vmax2(mx, b[k]); }
  bool p0, p1, p2, p3;
  wp::vibmax2(mn, pk2(lo), p0, p1);     // mn >= lo per half
  wp::vibmax2(pk2(hi), mx, p2, p3);     // hi >= mx per half
  return p0 && p1 && p2 && p3;
}

// Can this rectangle go through the packed path? Shape (32..256 rows, whole 8-column groups, not the forced origin
// cell unless run_generic set it up) and value range of its inputs (pk_bounds). Nothing is modified.
template <bool TALL = false>
BA_DEV bool pk_rect_ok(const Params& P, const RectArgs& a, bool origin_ok = false) {
  const int lane = wp::lane_id();
  const int H = a.H, W = a.W;
  const bool tall = BA_PK_TALL && TALL && H > 256 && (H & 255) == 0 && 2 * W <= (int)P.max_size;      // swept in 256-row chunks (place_rect_pk_tall)
  if (!(H == 32 || H == 64 || H == 128 || H == 256 || tall) || W <= 0 || (W & 7) || W > (tall ? 4096 : 256) || a.ncols <= 0) return false;
  if (a.vec_base == 0 && a.col_base == 0 && !origin_ok) return false;   // forced origin cell (scan_block.rs:1130-1132): only as set up by run_generic
  int GL, GH;
  pk_bounds(W, P.gap_open, P.gap_extend, P.pk_smax, GL, GH);
  // raw border values v become v + off_add (saturating in the reference): exact and inside [GL, GH] iff
  // v is inside [GL - off_add, GH - off_add] (clipped to i16)
  const int lo_b = wp::imax(GL - a.off_add, kI16Min), hi_b = wp::imin(GH - a.off_add, kI16Max);
  // every lane tests 8 consecutive entries of both input borders (entries >= H: lanes mirror the first ones)
  bool ok = lo_b <= hi_b && a.corner >= 0 && a.corner <= GH;
  for (int e0 = 0; e0 < ((BA_PK_TALL && TALL) ? H : 1); e0 += 256) {
    const int e = e0 + ((8 * lane) & (((BA_PK_TALL && TALL) ? wp::imin(H, 256) : H) - 1));
    const uint4 d = *(const uint4*)(a.AD + e), c = *(const uint4*)(a.AC + e);
    const uint32_t D[4] = {d.x, d.y, d.z, d.w}, C[4] = {c.x, c.y, c.z, c.w};
    ok = ok && pk_in_range<4>(D, C, lo_b, hi_b);
  }
  return wp::ballot(!ok) == 0u;
}

// Packed replacement of place_rect for sequence-sequence rectangles with borders in shared memory (generic
// phase, one alignment per warp; lanes >= G mirror lanes < G). Only for rectangles pk_rect_ok accepted.
// TRACE: a.tw receives H / 8 words per column (Rect layout 3, see pk_cols8); TRACE keeps four registers per lane (its
// trace words hold eight nibbles per lane), so rectangles below 256 rows use G < 32 lanes there. Without TRACE a
// rectangle of 128 / 64 / 32 rows runs with 2 / 1 / 1 registers per lane on 32 / 32 / 16 lanes.
template <int KIND, bool XDROP, bool TRACE, int K, int GT = 0>
BA_DEV void place_rect_pk_k(const unsigned char* smem, const Params& P, const PkConst& kc, const uint8_t* vec, const uint8_t* col,
                            const RectArgs& a, uint32_t* fr, int& bv, unsigned& bkey) {
  const int lane = wp::lane_id();
  const int H = a.H, W = a.W;
  const int G = GT ? GT : H / (2 * K);
  const int lg = lane & (G - 1);
  uint32_t D[K], C[K];
  pk_loadk<K>(a.AD, lg, G, D);
  pk_loadk<K>(a.AC, lg, G, C);
  bv = 0; bkey = 15u << kKeyClsShift;
  const uint32_t oa2 = pk2(a.off_add);
#pragma unroll
  for (int k = 0; k < K; k++) { D[k] = wp::vadd2(D[k], oa2); C[k] = wp::vadd2(C[k], oa2); }

  PkScorer<KIND> sc;
  sc.init(smem, P);
  {
    const uint8_t* vl = vec + a.vec_base + K * lg;
    const uint8_t* vh = vl + K * G;
    if (K == 4) sc.rows(*(const uint32_t*)vl, *(const uint32_t*)vh);
    else if (K == 2) sc.rows(*(const uint16_t*)vl, *(const uint16_t*)vh);
    else sc.rows(*vl, *vh);
  }
  uint32_t m[K], mc[PkMc<TRACE, K>::kN] = {};
#pragma unroll
  for (int k = 0; k < K; k++) m[k] = 0u;
  const bool writer = lane == G - 1;
  for (int cb = 0; cb < a.ncols; cb += 8) {      // a.ncols < W: early break (run_generic), the remaining columns do not exist
    const uint2 cw = *(const uint2*)(col + a.col_base + cb);
    const int n8 = wp::imin(8, a.ncols - cb);
    pk_cols8<KIND, XDROP, 5, TRACE, K, GT>(sc, kc, G, lg, cw.x, cw.y, D, C, (uint32_t)a.corner & 0xffffu, cb, m, mc, fr, writer, a.tw, lane < G, n8);
    wp::syncwarp();
    if (lane < n8) { const uint32_t v = fr[lane]; a.OD[cb + lane] = (int16_t)(v >> 16); a.OR_[cb + lane] = (int16_t)(v & 0xffffu); }
    wp::syncwarp();
  }
  if (lane < G) { pk_storek<K>(a.AD, lg, G, D); pk_storek<K>(a.AC, lg, G, C); }
  if (XDROP) {
    bv = pk_lane_max<K>(m);
    if (wp::red_max(bv) > a.key_thr) bkey = pk_lane_key<K, PkMc<TRACE, K>::kN>(m, mc, lg, G, bv);   // warp-uniform
  }
  else bv = wp::imax(bv, wp::imax(wp::h_lo(m[0]), wp::h_hi(m[0])));
  wp::syncwarp();
}
// Rectangles taller than 256 rows (blocks 512 .. 16384: the grow rectangles of long alignments) as a stack of 256-row
// chunks, chunk-major like the exact path: chunk ch sweeps all columns, reading the bottom row of chunk ch - 1 from
// a.OD / a.OR_ (which it then overwrites with its own, so that the last chunk leaves the rectangle's bottom row there)
// and, for TRACE, that row's "opened here" bits from `ecarry`. Trace words: layout 2 (H / 8 words per column, 32 per chunk).
template <int KIND, bool XDROP, bool TRACE>
BA_DEV void place_rect_pk_tall(const unsigned char* smem, const Params& P, const PkConst& kc, const uint8_t* vec, const uint8_t* col,
                               const RectArgs& a, uint32_t* fr, int& bv, unsigned& bkey, uint8_t* ecarry) {
  const int lane = wp::lane_id();
  const int H = a.H;
  // `ecarry` (max_size bytes) in two halves, written and read by alternate chunks: a rectangle this tall is at most
  // max_size / 2 columns wide (grow rectangles) or 8 (shift steps)
  const int ehalf = (int)(P.max_size >> 1);
  constexpr int G = 32;
  bv = 0; bkey = 15u << kKeyClsShift;
  const uint32_t oa2 = pk2(a.off_add);
  PkScorer<KIND> sc;
  sc.init(smem, P);
  uint32_t* tops = fr + 8;                 // staging of the row above (fr[0..8) is the column routine's own bottom row)
  uint32_t corner = (uint32_t)a.corner & 0xffffu;
  for (int ch = 0; ch * 256 < H; ch++) {
    uint32_t D[4], C[4];
    pk_loadk<4>(a.AD + 256 * ch, lane, G, D);
    pk_loadk<4>(a.AC + 256 * ch, lane, G, C);
#pragma unroll
    for (int k = 0; k < 4; k++) { D[k] = wp::vadd2(D[k], oa2); C[k] = wp::vadd2(C[k], oa2); }
    // the old border value of this chunk's last row is the next chunk's corner (the store below overwrites it)
    const uint32_t next_corner = (uint32_t)wp::shfl_idx((int)D[3], G - 1) >> 16;
    sc.rows(*(const uint32_t*)(vec + a.vec_base + 256 * ch + 4 * lane), *(const uint32_t*)(vec + a.vec_base + 256 * ch + 128 + 4 * lane));
    uint32_t m[4] = {0u, 0u, 0u, 0u}, mc[PkMc<TRACE, 4>::kN] = {};
    uint32_t dprev = corner;
    for (int cb = 0; cb < a.ncols; cb += 8) {
      const uint2 cw = *(const uint2*)(col + a.col_base + cb);
      const int n8 = wp::imin(8, a.ncols - cb);
      // the row above: the bottom row of chunk ch - 1; for the top chunk R = open - extend (U = 0: cannot win) and D = MIN = 0
      if (lane < 8) tops[lane] = ch > 0 ? ((uint32_t)(uint16_t)a.OR_[cb + lane] | ((uint32_t)(uint16_t)a.OD[cb + lane] << 16))
                                        : (uint32_t)(uint16_t)kc.or2;
      wp::syncwarp();
      pk_cols8<KIND, XDROP, 5, TRACE, 4, 32, true>(sc, kc, G, lane, cw.x, cw.y, D, C, 0u, cb, m, mc, fr, lane == G - 1,
                                                    a.tw ? a.tw + 32 * ch : nullptr, TRACE, n8, tops, dprev,
                                                    ch > 0 ? ecarry + ((ch + 1) & 1) * ehalf + cb : nullptr, ecarry + (ch & 1) * ehalf + cb, H >> 3);
      dprev = tops[7] >> 16;       // (the top chunk: 0, the diagonal input after column 0 is MIN)
      wp::syncwarp();
      if (lane < n8) { const uint32_t v = fr[lane]; a.OD[cb + lane] = (int16_t)(v >> 16); a.OR_[cb + lane] = (int16_t)(v & 0xffffu); }
      wp::syncwarp();
    }
    pk_storek<4>(a.AD + 256 * ch, lane, G, D);
    pk_storek<4>(a.AC + 256 * ch, lane, G, C);
    corner = next_corner;
    if (XDROP) {
      // per-chunk maximum and key (row = 256 ch + row in chunk; 256 is a multiple of 16, the lane classes do not move)
      const int cbv = pk_lane_max<4>(m);
      if (wp::red_max(cbv) > a.key_thr) {
        const unsigned ck = pk_lane_key<4, PkMc<TRACE, 4>::kN>(m, mc, lane, G, cbv) + (unsigned)(256 * ch);
        if (cbv > bv || (cbv == bv && ck > bkey)) bkey = ck;
      }
      bv = wp::imax(bv, cbv);
    } else {
      bv = wp::imax(bv, wp::imax(wp::h_lo(m[0]), wp::h_hi(m[0])));
    }
    wp::syncwarp();
  }
}

template <int KIND, bool XDROP, bool TRACE, bool TALL = false>
BA_DEV void place_rect_pk(const unsigned char* smem, const Params& P, const PkConst& kc, const uint8_t* vec, const uint8_t* col,
                          const RectArgs& a, uint32_t* fr, int& bv, unsigned& bkey, uint8_t* ecarry = nullptr) {
  if (BA_PK_TALL && TALL && a.H > 256) { place_rect_pk_tall<KIND, XDROP, TRACE>(smem, P, kc, vec, col, a, fr, bv, bkey, ecarry); return; }
  constexpr int K128 = (TRACE || !(BA_PK_KVAR & 2)) ? 4 : 2, K64 = (TRACE || !(BA_PK_KVAR & 1)) ? 4 : 1;
  // (one call site per distinct K: every call is an inlined copy of the rectangle code)
  // With both variants on, 128- and 256-row rectangles always use all 32 lanes: their group size is a compile-time 32.
  constexpr int G4 = (K128 != 4 && K64 != 4 && BA_PK_GT) ? 32 : 0, G2 = BA_PK_GT ? 32 : 0;
  if (K128 != 4 && a.H == 128) place_rect_pk_k<KIND, XDROP, TRACE, K128, G2>(smem, P, kc, vec, col, a, fr, bv, bkey);
#if BA_PK_GT > 1   // 64-row rectangles with a compile-time group size of their own (one more copy of the rectangle code)
  else if (K64 != 4 && a.H == 64) place_rect_pk_k<KIND, XDROP, TRACE, K64, 32>(smem, P, kc, vec, col, a, fr, bv, bkey);
#endif
  else if (K64 != 4 && a.H <= 64) place_rect_pk_k<KIND, XDROP, TRACE, K64>(smem, P, kc, vec, col, a, fr, bv, bkey);
  else place_rect_pk_k<KIND, XDROP, TRACE, 4, G4>(smem, P, kc, vec, col, a, fr, bv, bkey);
}

}  // namespace ba
