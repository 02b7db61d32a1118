// ba_types.h -- plain structs shared by the host runtime (ba_runtime.cu) and the device code
// (ba_kernel.cuh). No CUDA or torch types here.
#pragma once
#include <stdint.h>

namespace ba {

// Constants of the reference's AVX2 build that leak into observable behaviour
// (reference: src/avx2.rs:11-16, src/scan_block.rs:787-790).
constexpr int kL = 16;             // AVX2 lanes: min block size clamp, argmax tie-break class, scan phantoms
constexpr int kZero = 1 << 14;     // stored i16 v  <->  true score off + v - kZero
constexpr int kStep = 8;
constexpr int kXDropIter = 2;
// Largest block size served: what percent_len can return (src/lib.rs:109-111). The reference itself accepts sizes below
// 65535 (src/scan_block.rs:855), i.e. also 32768, which it does not recommend; 32768 rows + 16384 columns + the lane class
// no longer fit the 32-bit argmax key below.
constexpr int kMaxBlock = 16384;
// X-drop argmax key (total order of src/scan_block.rs:1194-1201, src/avx2.rs:271-274): (15 - lane class) | column + 1 | row
constexpr int kKeyClsShift = 28, kKeyColShift = 14;
constexpr unsigned kKeyRowMask = 0x3fffu;

// Block<TRACE, X_DROP, ...> const generics as bit flags (reference: src/scan_block.rs:89)
enum Flags : int { kTrace = 1, kXDrop = 2, kLocalStart = 4, kFreeQueryStartGaps = 8, kFreeQueryEndGaps = 16 };
// template-only flag: the instantiation consults Params::ext_flags (LOCAL_START / FREE_QUERY_START_GAPS) at run time
constexpr int kExt = 4;
// scoring kinds
enum Scoring : int { kNuc = 0, kAA = 1, kByte = 2, kProfile = 3 };
enum Dir : int { kRight = 0, kDown = 1, kGrow = 2 };

// per-pair status written by the kernel
enum Status : uint32_t { kOk = 0, kNotRun = 1, kTraceOverflow = 2, kRectOverflow = 3, kCigarOverflow = 4, kTraceGone = 5 };

// Device-side view of one AAProfile (reference: src/scores.rs:454-468). `pos_aa` is the only score
// layout kept on the device (row per position, 32 i8 each); the reference's transposed `aa_pos`
// copy holds the same numbers.
struct ProfileDev {
  const int8_t* pos_aa;          // [curr_len][32]
  const int16_t* gap_open_C;     // [curr_len]
  const int16_t* gap_close_C;    // [curr_len]
  const int16_t* gap_open_R;     // [curr_len]
  uint32_t len;                  // str_len
  int32_t gap_extend;
  // derived layouts for the packed fast phase (built on the device next to the arrays above; null when not built):
  // tp: the scores transposed, row per residue, [32][tlen] i8 (the numbers of the reference's aa_pos, scores.rs:457):
  //     the 4 / 8 consecutive positions a step needs for one residue are one 32-bit / 64-bit load;
  // gp: per position {gap_open_C + gap_extend, gap_close_C, gap_open_R} as 32-bit "v * 65537" operands of the packed
  //     recurrence, 16 bytes per position = one load per rectangle column
  const int8_t* tp; const uint32_t* gp; uint32_t tlen; uint32_t pad_;
};

// Construction of one padded device profile (ba_profile_build_kernel): AAProfile::new's -128 defaults
// (src/scores.rs:473-486) everywhere, then either the leading `np` positions of a host AAProfile (kind 0) or
// AAProfile::set_all / set_all_rev + the gap setters (src/scores.rs:532-538, 548-580, 676-715) applied to a raw
// position-major score table (kind 1).
struct ProfBuild {
  uint64_t src_off;      // kind 0: byte offset of the staged compact profile; kind 1: offset of its scores [len][order_len]
  uint64_t gap_src_off;  // kind 1: offset into the per-position gap arrays (len + 1 entries per profile)
  uint64_t dst_off;      // byte offset inside the profile arena (64-byte aligned)
  uint32_t np;           // kind 0: staged positions; kind 1: len
  uint32_t curr_len;     // positions of the padded device profile
  int32_t gap_extend;
  uint32_t derive;       // also build the derived layouts (ProfileDev::tp / gp) behind the base arrays
};
struct ProfBuildArgs {
  const ProfBuild* desc; uint32_t n; int32_t kind;
  const uint8_t* stage;  // kind 0
  uint8_t* arena;
  const int8_t* scores; const int8_t *gap_open_C, *gap_close_C, *gap_open_R;   // kind 1; gap arrays may be null
  int32_t all_open_C, all_close_C, all_open_R;
  uint32_t order_len, left_shift, right_shift, rev;
  int8_t inv[32];        // residue code -> index in `order` (last occurrence wins like the reference's loop), -1 = absent
  uint32_t* err;         // set when a gap open cost is not negative
};

// one computed rectangle (reference: Trace::add_block, src/scan_block.rs:1428-1443)
struct Rect {
  uint32_t row, col;             // top-left cell in DP-matrix coordinates
  uint16_t h, w;                 // h = extent along the vector direction, w = number of sequential columns
  uint32_t right;                // 1: vectors run along rows (query); 0: along columns (reference)
  uint32_t word_off;             // first trace word of this rectangle in the warp's arena
};

struct DevResult {
  int32_t score;
  uint32_t query_idx, reference_idx;
  uint32_t status;
  uint64_t cells;                // sum of h*w over every rectangle handed to place_rect
  uint32_t steps;
  uint32_t cigar_n;              // number of runs written for this pair
  uint64_t cigar_off;            // offset (in runs) into the cigar stream
  uint32_t rect_n;               // rectangles on the trace stack at the end (TRACE)
  uint32_t warp;                 // trace arena that holds this pair's trace (TRACE)
};

// one record per step, debug builds only (mirrors ora_step in oracle/ba_oracle.h)
struct StepLog { int32_t dir; uint32_t i, j, block_size; int32_t off; int16_t max, right_max, down_max; };

// Uniform constants of the packed recurrence (ba_packed.cuh), filled in by the host so that the kernel reads them
// as constant-bank operands instead of keeping (or recomputing) them in registers.
struct PkConst {
  uint32_t ge2, go1, or2;   // packed gap_extend; gap_open * 65537 (one 32-bit add = packed add, see pk_cols8); packed open - extend
  uint32_t kge[4];          // packed (k + 1) * gap_extend
  uint32_t dec[5];          // Kogge-Stone decays: packed (4 << s) * gap_extend
  uint32_t lane1;           // 4 * gap_extend * 65537: lane lg's packed decay 4 * lg * gap_extend = lg * lane1 + (lg ? 65536 : 0)
  uint32_t ge1;             // gap_extend * 65537: the same decay for K rows per lane = (K * lg) * ge1 + (lg ? 65536 : 0)
};

struct Params {
  uint32_t n_pairs;
  const uint32_t* order;         // optional processing order (pair ids), longest first
  const uint8_t* seq;            // arena of padded, converted sequences (pad byte at index 0 of each)
  const uint64_t* q_off; const uint32_t* q_len;
  const uint64_t* r_off; const uint32_t* r_len;
  const ProfileDev* profiles;    // kProfile: one per pair (r_off/r_len unused)
  const int8_t* matrix;          // kNuc: 128 B, kAA: 864 B, kByte: {match, mismatch}
  int32_t gap_open, gap_extend;
  uint32_t min_size, max_size;   // already clamped to >= kL, powers of two
  int32_t x_drop;
  int32_t flags, scoring;
  int32_t pk_smax;               // max(0, largest matrix entry): growth bound of the packed path's range guard
  uint32_t pk_enable;            // packed 2 x i16 DP path on (sequence-sequence)
  PkConst kc;
  uint32_t pk_fast;              // the fast phase is the packed one (pk_fast_step): entering it needs pk_borders_ok
  uint32_t ext_flags;            // kLocalStart | kFreeQueryStartGaps (only honoured by kernels instantiated with kExt)
  uint32_t* trace_zwords;        // zero-mask words per slot (TRACE && LOCAL_START), same stride as trace_words
  DevResult* out;
  uint32_t* ticket;              // work counter
  // per-warp scratch in global memory
  // a "slot" is one alignment in flight: 1 per warp, or 4 / 8 per warp when the packed fast phase is on
  uint32_t slots_per_warp;
  uint32_t fast_block;           // 0: fast phase off; 32 / 64: block size served by the fast phase (== min_size)
  int16_t* ckpt;                 // 4 * max(max_size, 32) int16 per slot: checkpoint borders
  int16_t* gborders;             // kernels with global live borders (FM >= 32): 4 * max(max_size, 32) int16 per warp
  // TRACE: every pair of a launch owns one trace arena (words + rectangle records + run scratch), arena_of[pair] = its
  // index. The trace stays resident after the alignment, so the CIGARs are produced by a second kernel with one walk per
  // lane (ba_traceback_batch_kernel) instead of one walk per warp inside the alignment kernel.
  const uint32_t* arena_of;
  uint32_t* trace_words; uint64_t trace_words_per_warp;   // per arena (names kept from the per-warp arenas of v0)
  // shared overflow pool for rectangles that do not fit a slot's own arena: bump-allocated in units of 16 words
  uint32_t* trace_pool; uint32_t* trace_pool_cursor; uint64_t trace_pool_units;
  Rect* rects; uint32_t rects_per_warp;                   // per arena
  uint32_t* run_scratch; uint32_t runs_per_warp;  // per arena: reversed CIGAR runs while walking back
  // cigar output stream: runs packed as (len << 4) | op, allocated with atomicAdd on *cigar_used
  uint32_t* cigar_stream; uint64_t cigar_cap; unsigned long long* cigar_used;
  uint32_t cigar_eq;
  // pairs whose trace arena overflowed (they are re-run with worst-case arenas)
  uint32_t* overflow_list; uint32_t* overflow_n;

  // debug
  StepLog* step_log; uint32_t step_log_cap; uint32_t* step_log_n;   // only honoured for n_pairs == 1
};

}  // namespace ba
