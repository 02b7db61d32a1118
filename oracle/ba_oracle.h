/* ba_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, never linked by the product).
 * See ba_oracle.cpp for what each entry point restates (reference file:line). */
#ifndef BA_ORACLE_H
#define BA_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OraBlock OraBlock;
typedef struct OraProfile OraProfile;

/* scan_block.rs:1887-1893 */
typedef struct ora_result { int32_t score; size_t query_idx; size_t reference_idx; } ora_result;
/* cigar.rs:36-39 */
typedef struct ora_oplen { uint8_t op; size_t len; } ora_oplen;
/* one record per iteration of the step loop (scan_block.rs:130), for step-by-step diffing */
typedef struct ora_step {
  int32_t dir;            /* 0 right, 1 down, 2 grow */
  uint32_t i, j, block_size;
  int32_t off;
  int16_t max, right_max, down_max;
} ora_step;

enum { ORA_MATRIX_NUC = 0, ORA_MATRIX_AA = 1, ORA_MATRIX_BYTE = 2 };
/* Block<TRACE, X_DROP, LOCAL_START, FREE_QUERY_START_GAPS, FREE_QUERY_END_GAPS> as bit flags */
enum { ORA_TRACE = 1, ORA_XDROP = 2, ORA_LOCAL_START = 4, ORA_FREE_QUERY_START_GAPS = 8, ORA_FREE_QUERY_END_GAPS = 16 };
enum { ORA_OK = 0, ORA_ERR_GAPS = 1, ORA_ERR_SIZE = 2, ORA_ERR_XDROP = 3, ORA_ERR_BAD_FLAGS = 4,
       ORA_ERR_TOO_LONG = 5, ORA_ERR_CHAR = 6 };

OraBlock* ora_block_new(size_t query_len, size_t reference_len, size_t max_size, int flags);
void ora_block_free(OraBlock*);
int ora_align(OraBlock*, const uint8_t* q_padded, size_t qlen, const uint8_t* r_padded, size_t rlen,
              int matrix_kind, const int8_t* matrix, int8_t gap_open, int8_t gap_extend,
              size_t min_size, size_t max_size, int32_t x_drop);
int ora_align_profile(OraBlock*, const uint8_t* q_padded, size_t qlen, const OraProfile*,
                      size_t min_size, size_t max_size, int32_t x_drop);
ora_result ora_res(const OraBlock*);
uint64_t ora_cells(const OraBlock*);
uint64_t ora_steps(const OraBlock*);
long ora_cigar(OraBlock*, size_t query_idx, size_t reference_idx, int eq, const uint8_t* q_padded, size_t qlen,
               const uint8_t* r_padded, size_t rlen, ora_oplen* out, size_t cap);
void ora_enable_step_log(OraBlock*, int on);
size_t ora_align_logged_steps(OraBlock*, ora_step* out, size_t cap);

OraProfile* ora_profile_new(size_t str_len, size_t block_size, int8_t gap_extend);
OraProfile* ora_profile_from_bytes(const uint8_t* b, size_t len, size_t block_size, int8_t match, int8_t mismatch,
                                   int8_t gap_open_C, int8_t gap_close_C, int8_t gap_open_R, int8_t gap_extend);
void ora_profile_free(OraProfile*);
size_t ora_profile_len(const OraProfile*);
int ora_profile_clear(OraProfile*, size_t str_len, size_t block_size);
int ora_profile_set(OraProfile*, size_t i, uint8_t b, int8_t score);
int ora_profile_set_all(OraProfile*, const uint8_t* order, size_t order_len, const int8_t* scores,
                        size_t scores_len, size_t left_shift, size_t right_shift, int rev);
int ora_profile_set_gap_open_C(OraProfile*, size_t i, int8_t g);
int ora_profile_set_gap_close_C(OraProfile*, size_t i, int8_t g);
int ora_profile_set_gap_open_R(OraProfile*, size_t i, int8_t g);
int ora_profile_set_all_gap_open_C(OraProfile*, int8_t g);
int ora_profile_set_all_gap_close_C(OraProfile*, int8_t g);
int ora_profile_set_all_gap_open_R(OraProfile*, int8_t g);
int8_t ora_profile_get(const OraProfile*, size_t i, uint8_t b);
int8_t ora_profile_get_gap_extend(const OraProfile*);
const int8_t* ora_profile_pos_aa(const OraProfile*);
const int16_t* ora_profile_gap_open_C(const OraProfile*);
const int16_t* ora_profile_gap_close_C(const OraProfile*);
const int16_t* ora_profile_gap_open_R(const OraProfile*);
size_t ora_profile_curr_len(const OraProfile*);

void ora_nuc_matrix_simple(int8_t match, int8_t mismatch, int8_t* out128);
void ora_aa_matrix_simple(int8_t match, int8_t mismatch, int8_t* out864);
int ora_pad(int matrix_kind, const uint8_t* s, size_t len, size_t block_size, int rev, uint8_t* out);
void ora_prefix_scan(const int16_t* in16, int16_t gap, int16_t* out16);
void ora_prefix_scan_consts(int16_t gap, int16_t* gap_all16, int16_t* consts16);

/* batch driver = the CPU baseline */
typedef struct ora_batch {
  size_t n;
  const uint8_t* arena;            /* padded sequences (pad byte first) */
  const uint64_t* q_off; const uint32_t* q_len;
  const uint64_t* r_off; const uint32_t* r_len;   /* unused when profiles != NULL */
  const OraProfile* const* profiles;              /* NULL for sequence-sequence */
  int matrix_kind; const int8_t* matrix;
  int8_t gap_open, gap_extend;
  size_t min_size, max_size;
  int32_t x_drop;
  int flags;
  int cigar_eq;
} ora_batch;
int ora_batch_align(const ora_batch*, int n_threads, ora_result* out, uint64_t* out_cells,
                    ora_oplen* cigar_arena, const uint64_t* cigar_off, uint32_t* cigar_len);
int ora_pad_batch(int matrix_kind, const uint8_t* raw, const uint64_t* off, size_t n, size_t block_size, uint8_t* out,
                  uint64_t* out_off, uint32_t* out_len);
double ora_last_batch_seconds(void);
int ora_hw_threads(void);

#ifdef __cplusplus
}
#endif
#endif
