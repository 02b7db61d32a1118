// ba_host.h -- host-side mirrors of the reference's scoring / string / CIGAR types
// (src/scores.rs, src/scan_block.rs:1790-1884, src/cigar.rs). These are the concrete definitions
// of the opaque structs declared in include/block_aligner_b200.h. They only hold data and do
// host-side bookkeeping; every alignment runs on the GPU.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

#include "ba_types.h"

// scores.rs:40-44: `#[repr(C, align(32))] struct AAMatrix { scores: [i8; 27 * 32] }`
struct alignas(32) AAMatrix { int8_t scores[27 * 32]; };
// scores.rs:142-146: `#[repr(C, align(32))] struct NucMatrix { scores: [i8; 8 * 16] }`
struct alignas(32) NucMatrix { int8_t scores[8 * 16]; };

// scores.rs:454-468. `aa_pos` (the transposed i16 copy) is not kept: it holds the same numbers as
// `pos_aa` and the device only needs one layout.
struct AAProfile {
  std::vector<int8_t> pos_aa;                 // [max_len][32]
  std::vector<int16_t> gap_open_C, gap_close_C, gap_open_R;  // [max_len]
  int8_t gap_extend;
  size_t max_len, curr_len, str_len;
  size_t hi_pos = 0;   // highest position written since new / clear: positions above it still hold the -128 defaults
};

// scan_block.rs:1790-1793
struct PaddedBytes {
  std::vector<uint8_t> s;
  size_t len;
};

// cigar.rs:36-45 (OpLen is declared in the public header)
struct CigarRun { uint8_t op; size_t len; };
struct Cigar {
  std::vector<CigarRun> s;   // s[0] is the sentinel; runs are stored reversed like the reference
  size_t idx;
};

namespace ba { namespace host {

inline uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
// Matrix::NULL after convert_char (scores.rs:83, 168, 239)
inline uint8_t null_code(int scoring) { return scoring == kNuc ? (uint8_t)'Z' : (scoring == kByte ? (uint8_t)0 : (uint8_t)26); }
// Matrix::convert_char (scores.rs:130-134, 212-216, 270-272)
inline uint8_t convert_char(int scoring, uint8_t c, bool* ok) {
  *ok = true;
  if (scoring == kByte) return c;
  c = upper(c);
  if (scoring == kNuc) { if (!(c >= 'A' && c <= 'Z')) *ok = false; return c; }
  if (!(c >= 'A' && c <= 'A' + 26)) *ok = false;
  return (uint8_t)(c - 'A');
}

size_t profile_len(const AAProfile* p);
size_t profile_curr_len(const AAProfile* p);
int profile_gap_extend(const AAProfile* p);
// number of leading positions that may differ from the defaults (the rest of the padded profile is -128 everywhere)
size_t profile_used_len(const AAProfile* p);
// first `np` positions of every array
void profile_export(const AAProfile* p, size_t np, int8_t* pos_aa, int16_t* open_C, int16_t* close_C, int16_t* open_R);

}}  // namespace ba::host
