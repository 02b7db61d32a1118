/* c_example.c -- the three usage patterns of the reference's C API (global alignment, global alignment
 * with traceback, sequence-to-profile; cf. the reference's c/example.c), written against
 * include/block_aligner_b200.h. Prints one line per example; tests/test_c_api.py checks the numbers
 * against the CPU oracle. Build: see tests/test_c_api.py. */
#include <stdio.h>
#include <string.h>

#include "block_aligner_b200.h"

static void seq_seq(void) {
  const char* a_str = "AAAAAAAA";
  const char* b_str = "AARAAAA";
  size_t a_len = strlen(a_str), b_len = strlen(b_str);
  SizeRange range = {32, 32};
  Gaps gaps = {-11, -1};
  PaddedBytes* a = block_new_padded_aa(a_len, range.max);
  PaddedBytes* b = block_new_padded_aa(b_len, range.max);
  block_set_bytes_padded_aa(a, (const uint8_t*)a_str, a_len, range.max);
  block_set_bytes_padded_aa(b, (const uint8_t*)b_str, b_len, range.max);
  BlockHandle block = block_new_aa(a_len, b_len, range.max);
  block_align_aa(block, a, b, &BLOSUM62, gaps, range, 0);
  AlignResult res = block_res_aa(block);
  printf("global score=%d idx=(%lu,%lu)\n", res.score, (unsigned long)res.query_idx, (unsigned long)res.reference_idx);
  block_free_aa(block);
  block_free_padded_aa(a);
  block_free_padded_aa(b);
}

static void seq_seq_trace(void) {
  const char* a_str = "AAAAAAAA";
  const char* b_str = "AARAAAA";
  size_t a_len = strlen(a_str), b_len = strlen(b_str);
  SizeRange range = {32, 32};
  Gaps gaps = {-11, -1};
  PaddedBytes* a = block_new_padded_aa(a_len, range.max);
  PaddedBytes* b = block_new_padded_aa(b_len, range.max);
  block_set_bytes_padded_aa(a, (const uint8_t*)a_str, a_len, range.max);
  block_set_bytes_padded_aa(b, (const uint8_t*)b_str, b_len, range.max);
  BlockHandle block = block_new_aa_trace(a_len, b_len, range.max);
  block_align_aa_trace(block, a, b, &BLOSUM62, gaps, range, 0);
  AlignResult res = block_res_aa_trace(block);
  Cigar* cigar = block_new_cigar(res.query_idx, res.reference_idx);
  block_cigar_aa_trace(block, res.query_idx, res.reference_idx, cigar);
  printf("trace score=%d idx=(%lu,%lu) cigar=", res.score, (unsigned long)res.query_idx, (unsigned long)res.reference_idx);
  const char ops[] = {' ', 'M', '=', 'X', 'I', 'D'};
  for (size_t i = 0; i < block_len_cigar(cigar); i++) {
    OpLen o = block_get_cigar(cigar, i);
    printf("%lu%c", (unsigned long)o.len, ops[o.op]);
  }
  printf("\n");
  block_free_cigar(cigar);
  block_free_aa_trace(block);
  block_free_padded_aa(a);
  block_free_padded_aa(b);
}

static void seq_profile(void) {
  const char* a_str = "AAAAAAAA";
  size_t a_len = strlen(a_str), b_len = 7;
  SizeRange range = {32, 32};
  PaddedBytes* a = block_new_padded_aa(a_len, range.max);
  block_set_bytes_padded_aa(a, (const uint8_t*)a_str, a_len, range.max);
  AAProfile* b = block_new_aaprofile(b_len, range.max, -1);
  for (size_t i = 1; i <= b_len; i++)
    for (int c = 'A'; c <= 'Z'; c++) block_set_aaprofile(b, i, (uint8_t)c, c == a_str[i - 1] ? 1 : -1);
  for (size_t i = 0; i < b_len; i++) {
    block_set_gap_open_C_aaprofile(b, i, -10);
    block_set_gap_close_C_aaprofile(b, i, 0);
    block_set_gap_open_R_aaprofile(b, i, -10);
  }
  BlockHandle block = block_new_aa(a_len, b_len, range.max);
  block_align_profile_aa(block, a, b, range, 0);
  AlignResult res = block_res_aa(block);
  printf("profile score=%d idx=(%lu,%lu)\n", res.score, (unsigned long)res.query_idx, (unsigned long)res.reference_idx);
  block_free_aa(block);
  block_free_padded_aa(a);
  block_free_aaprofile(b);
}

int main(void) {
  seq_seq();
  seq_seq_trace();
  seq_profile();
  return 0;
}
