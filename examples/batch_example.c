/* Batch API (include/block_aligner_b200.h, Part 2): many (query, reference) pairs in one call, CIGARs included.
 * The loop a user of the reference writes around Block::align (e.g. examples/nanopore_bench.rs:86-95) becomes one
 * ba_align_batch_cigar call.
 *
 *   gcc -O1 -I include examples/batch_example.c -o examples/batch_example \
 *       -L block_aligner_b200 -lblock_aligner_b200 -Wl,-rpath,$PWD/block_aligner_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "block_aligner_b200.h"

int main(void) {
  const char* queries[] = {"TTTTTTTTAAAAAAATTTTTTTTT", "AAAAAAAAATTGCGCT", "ACGTACGTACGTACGTACGTACGTACGTACGT"};
  const char* references[] = {"TTAAAAAAATTTTTTTTTTTT", "AAAAAAAAAGCGC", "ACGTACGTACGAACGTACGTACGTACGTACGT"};
  enum { N = 3 };

  /* concatenate into arenas + offset tables (n + 1 entries) */
  uint64_t q_off[N + 1] = {0}, r_off[N + 1] = {0};
  for (int k = 0; k < N; k++) { q_off[k + 1] = q_off[k] + strlen(queries[k]); r_off[k + 1] = r_off[k] + strlen(references[k]); }
  uint8_t* q = malloc(q_off[N] + 1), *r = malloc(r_off[N] + 1);
  for (int k = 0; k < N; k++) { memcpy(q + q_off[k], queries[k], strlen(queries[k])); memcpy(r + r_off[k], references[k], strlen(references[k])); }

  BaAligner* al = NULL;
  int rc = ba_create(0, &al);
  if (rc) { fprintf(stderr, "%s: %s\n", ba_error_string(rc), ba_last_error_message()); return 1; }

  BaConfig cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.scoring = BA_SCORING_NUC;
  cfg.flags = BA_TRACE;                 /* Block::<true, false> */
  cfg.matrix = &NW1;                    /* the crate's NW1 static */
  cfg.gaps.open = -2; cfg.gaps.extend = -1;
  cfg.size.min = 32; cfg.size.max = 256;
  cfg.x_drop = 0;
  cfg.cigar_eq = 1;                     /* '=' / 'X' instead of 'M' (Trace::cigar_eq) */

  AlignResult res[N];
  uint32_t runs[256], run_len[N];
  uint64_t run_off[N];
  size_t used = 0;
  BaStats st;
  rc = ba_align_batch_cigar(al, &cfg, N, q, q_off, r, r_off, res, runs, 256, run_off, run_len, &used, &st);
  if (rc) { fprintf(stderr, "%s: %s\n", ba_error_string(rc), ba_last_error_message()); return 1; }

  for (int k = 0; k < N; k++) {
    char cigar[128];
    ba_cigar_format(runs + run_off[k], run_len[k], cigar, sizeof(cigar));
    printf("pair %d: score=%d idx=(%zu,%zu) cigar=%s\n", k, res[k].score, (size_t)res[k].query_idx, (size_t)res[k].reference_idx, cigar);
  }
  printf("%llu cells in %u kernel launch(es)\n", (unsigned long long)st.cells, st.kernel_launches);
  ba_destroy(al);
  free(q); free(r);
  return 0;
}
