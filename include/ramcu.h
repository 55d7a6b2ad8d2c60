/*
 * ramcu.h -- the extern "C" ABI of the B200-native RAMExtend extension path.
 *
 * Host code stays C: the reference's main() (ram_extend.c:217-790) keeps calling
 *     int extend_alignment(int direction, struct coreAlignment*, int ****score,
 *                          struct sequenceLibrary*, char *master, int BANDWIDTH, int CAPPENALTY,
 *                          int MINIMPROVEMENT, int L, int N, struct scoringSystem*, FILE*);
 * (ram_extend.h:9-13). repeatafterme_b200/csrc/ramx_boundary.c provides a strong definition
 * of that symbol which flattens the reference's structs and calls the entry points below;
 * everything behind them is hand-written sm_100a CUDA (csrc/ramcu.cu). There is NO CPU fallback:
 * every entry point fails (negative return, ramcu_last_error()) when no CUDA device is usable.
 *
 * Plain C types only; no torch / C++ types cross this boundary.
 *
 * Reference interface each item replaces:
 *   ramcu_core_t         struct coreAlignment                  common.h:80-98
 *   ramcu_scoring_t      struct scoringSystem                  score_system.h:7-17
 *   ramcu_params_t       BANDWIDTH/CAPPENALTY/MINIMPROVEMENT/L args + WHEN_TO_STOP global
 *                                                              ram_extend.h:9-13, ram_extend.c:61
 *   ramcu_library_*      struct sequenceLibrary.sequence       sequence.h:36-46 (built by
 *                        loadSequenceSubsetMinimal, sequence.c:505-923)
 *   ramcu_extend         extend_alignment()                    ram_extend.c:859-1258
 *                        (which drives compute_nw_row(), bnw_extend.c:750-1048)
 *   ramcu_job_* / ramcu_batch_run
 *                        the same call, split into prepare / enqueue / fetch so that many
 *                        independent families (one RAMExtend process each in
 *                        util/extend-stk.pl:365-371) share one context and run concurrently
 *   ramcu_seed_extend_batch / ramcu_global_align
 *                        cons_seed_extend() / pair_seed_extend() / global_align()  bnw_extend.c:167-649
 *   ramcu_extend_paths   the -outmat branch of extend_alignment()           ram_extend.c:1115-1132
 *   ramcu_extend_verbose the VERBOSE >= 10 dumps of extend_alignment()       ram_extend.c:949-959, 992-1090, 1177-1214
 *   ramcu_debug_rows     compute_nw_row() as called by the reference's unit tests
 *                        (bnw_extend.c:1583,1600): forced consensus, band rows returned
 */
#ifndef RAMCU_H
#define RAMCU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAMCU_ABI_VERSION 3
#define RAMCU_NCODES 5            /* matrix columns: A,C,G,T,N (sequence.h:7-15; 99 -> 4)      */
#define RAMCU_MAX_BANDWIDTH 1023  /* 2*B+1 <= 2047 band cells: 32 lanes x 64 cells                */

/* error codes (all negative) */
#define RAMCU_OK 0
#define RAMCU_ERR_CUDA (-1)         /* CUDA runtime / no device                                */
#define RAMCU_ERR_ARG (-2)          /* invalid argument                                        */
#define RAMCU_ERR_UNSUPPORTED (-3)  /* outside what the kernels implement      */
#define RAMCU_ERR_CAPACITY (-4)     /* family does not fit the device in one pass              */

typedef struct ramcu_ctx ramcu_ctx;
typedef struct ramcu_library ramcu_library;
typedef struct ramcu_job ramcu_job;

/* One MSA instance == one struct coreAlignment (common.h:80-98), flattened. Positions index the
 * library's sequence[] array exactly as in the reference (0-based, closed). */
typedef struct
{
  int64_t left_seq_pos;
  int64_t right_seq_pos;
  int64_t lower_seq_bound;
  int64_t upper_seq_bound;
  int8_t orient;             /* 1 = reverse strand */
  int8_t left_extendable;
  int8_t right_extendable;
  int8_t pad_[5];
} ramcu_core_t;

/* matrix[consensus base 0..3][sequence code]: codes 0..3 = A,C,G,T, 4 = N (SYM_N and anything
 * that is not 0..3). Accepted range: (max |entry|, |gapopen| + |gapextn|) x (L + 2 * bandwidth + 3) <= 12 000 000
 * (job creation refuses more: the reference's -1000000000 floor could be crossed). Results are exact for every
 * accepted value; matrices whose entries all fit int16 -- every matrix of score_system.c -- run the fast rows,
 * others the exact per-cell path on every row (same results, several times slower). */
typedef struct
{
  int32_t matrix[4][RAMCU_NCODES];
  int32_t gapopen;
  int32_t gapextn;
} ramcu_scoring_t;

typedef struct
{
  int32_t bandwidth;       /* MAXOFFSET/BANDWIDTH: band is 2*bandwidth+1 cells               */
  int32_t cappenalty;
  int32_t minimprovement;
  int32_t L;               /* max extension length                                            */
  int32_t when_to_stop;    /* WHEN_TO_STOP global (ram_extend.c:61,249)                        */
  int32_t lanes_per_instance; /* 0 = auto; else force 1,2,4,8,16,32 (tuning/tests)            */
  int32_t concurrent_instances; /* hint: instances of OTHER families expected to share the GPU
                                 * with this one (ramcu_batch_run); 0 = runs alone            */
  int32_t cta_thread_limit;   /* 0 = auto; else only tilings compiled for this CTA limit, i.e.
                               * 65536 / limit registers per thread (tuning/tests)              */
} ramcu_params_t;

typedef struct
{
  int32_t extended_bp;     /* == return value of extend_alignment()                           */
  int32_t rows_processed;  /* column-loop iterations executed                                 */
  int32_t stop_row;        /* row_idx of the WHEN_TO_STOP break (for the VERBOSE>=3 line), -1 */
  int32_t hit_limit;       /* 1 => reference prints "WARNING: Extended sequence ... limit"    */
  int32_t max_score;       /* max_extension_score                                             */
  int32_t n_extendable;    /* instances extendable in this direction                          */
  float kernel_ms;         /* device time of the extension kernels (CUDA events), if measured  */
  int32_t lanes_per_instance; /* configuration the launch used                                */
  int32_t passes_per_thread;  /* 2: the family's right and left pass ran in the same threads; 1: one pass per
                               * thread (on its own, or next to the other pass's CTAs in the same launch)    */
} ramcu_result_t;

const char *ramcu_last_error(void);
int ramcu_abi_version(void);

int ramcu_init(int device, ramcu_ctx **out);
void ramcu_destroy(ramcu_ctx *ctx);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = context's own. */
int ramcu_set_stream(ramcu_ctx *ctx, void *cuda_stream);
int ramcu_device_info(ramcu_ctx *ctx, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor);

/* Measurement aid: INT32 ALU peak of the device in giga instruction-lanes per second (independent
 * fused add+max chains on every SM) -- the roofline denominator bench.py reports against. */
int ramcu_int32_peak(ramcu_ctx *ctx, double *gops_per_s, double *elapsed_ms);

/* Built-in scoring systems of score_system.c (getMatrix / getRepeatScoutMatrix). */
int ramcu_scoring_by_name(const char *name, int rs_match, int rs_mismatch, int rs_gap,
                          ramcu_scoring_t *out);

/* Upload seqLib->sequence (1 byte per base: 0..3, 99) from HOST memory. The bytes are packed
 * on the device to 2 bits/base + an N bitmask (the form the DP kernels read). */
int ramcu_library_upload(ramcu_ctx *ctx, const uint8_t *seq_bytes, uint64_t length,
                         ramcu_library **out);
/* Same, but seq_bytes is already a DEVICE pointer (no host<->device copy). */
int ramcu_library_from_device(ramcu_ctx *ctx, const uint8_t *d_seq_bytes, uint64_t length,
                              ramcu_library **out);
/* 2bit-native upload (the flank feed without the byte expansion of sequence.c:752-822): `blob`
 * holds packed DNA exactly as stored in a .2bit record (kentsrc/twoBitNew.c:571-582: 4 bases per
 * byte, first base in the top bits, T0 C1 A2 G3). Each window copies `len` bases starting at base
 * `src_base` of the record whose packed DNA starts at byte `src_byte` of the blob to library
 * positions [dst, dst+len); windows must tile [0, total_length) in increasing dst order (the
 * concatenation loadSequenceSubsetMinimal builds). N ranges (the record's N-blocks clipped to the
 * windows) are given in library coordinates. HOST pointers. */
typedef struct
{
  uint64_t dst;
  uint64_t len;
  uint64_t src_byte;
  uint64_t src_base;
} ramcu_window_t;
int ramcu_library_from_twobit(ramcu_ctx *ctx, const uint8_t *blob, uint64_t blob_bytes,
                              int n_windows, const ramcu_window_t *windows, int n_nranges,
                              const uint64_t *n_start, const uint64_t *n_len,
                              uint64_t total_length, ramcu_library **out);
/* Test aid: copy the packed device library back (lib2: 2 bits/base, 16 bases per word,
 * A0 C1 G2 T3; libn: N bitmask, 32 bases per word). Sizes: ceil(len/32)*2 and ceil(len/32) words. */
int ramcu_library_download(ramcu_ctx *ctx, ramcu_library *lib, uint32_t *lib2_out, uint32_t *libn_out);
int ramcu_library_free(ramcu_ctx *ctx, ramcu_library *lib);

/* One call == one call of the reference's extend_alignment() for `direction` (1 = right,
 * 0 = left). Outputs (HOST memory):
 *   master_out    : >= params->L bytes; master_out[r] = consensus base (0..3) chosen at row r,
 *                   valid for r < res->rows_processed (caller maps r to master[] as in
 *                   ram_extend.c:1092-1095)
 *   ext_len_out   : n_cores ints; new right/leftExtensionLen, 0 = leave unchanged
 *   score_add_out : n_cores ints; amount to add to coreAlignment.score
 */
int ramcu_extend(ramcu_ctx *ctx, ramcu_library *lib, int direction, int n_cores,
                 const ramcu_core_t *cores, const ramcu_scoring_t *scoring,
                 const ramcu_params_t *params, uint8_t *master_out, int32_t *ext_len_out,
                 int32_t *score_add_out, ramcu_result_t *res);

/* Split form. create: flatten + upload instance descriptors, allocate device state.
 * run: enqueue the kernels on the context stream (asynchronous, inputs resident in HBM).
 * fetch: wait + copy results to host. A job can be run repeatedly. */
int ramcu_job_create(ramcu_ctx *ctx, ramcu_library *lib, int direction, int n_cores,
                     const ramcu_core_t *cores, const ramcu_scoring_t *scoring,
                     const ramcu_params_t *params, ramcu_job **out);
int ramcu_job_run(ramcu_ctx *ctx, ramcu_job *job);
int ramcu_job_fetch(ramcu_ctx *ctx, ramcu_job *job, uint8_t *master_out, int32_t *ext_len_out,
                    int32_t *score_add_out, ramcu_result_t *res);
int ramcu_job_free(ramcu_ctx *ctx, ramcu_job *job);

/* Enqueue many independent jobs (families) so that they share the GPU concurrently. */
int ramcu_batch_run(ramcu_ctx *ctx, int n_jobs, ramcu_job **jobs);

/* Both passes of one family in one launch. Replaces the two extend_alignment() calls of main()
 * (ram_extend.c:424-427 right, :506-509 left) WHEN the overlap-avoidance step between them
 * (ram_extend.c:445-499) leaves every lower/upperSeqBound unchanged: both passes then read the
 * same `cores`. The caller owns that check (csrc/ramx_boundary.c compares the bounds main() hands
 * to the left pass with the ones used here and re-runs the left pass if any differs).
 * CTAs of the two passes share the SMs, so one pass computes while the other waits for its
 * per-column exchange; results are bit-identical to two ramcu_extend() calls on the same cores.
 * If the two passes cannot be co-resident they run one after the other.
 * ramcu_job_create_pair + ramcu_batch_run(ctx, 2, {right, left}) + ramcu_job_fetch is the split
 * form; kernel_ms of either result is the device time of the whole pair. */
int ramcu_job_create_pair(ramcu_ctx *ctx, ramcu_library *lib, int n_cores, const ramcu_core_t *cores,
                          const ramcu_scoring_t *scoring, const ramcu_params_t *params,
                          ramcu_job **right, ramcu_job **left);
int ramcu_extend_pair(ramcu_ctx *ctx, ramcu_library *lib, int n_cores, const ramcu_core_t *cores,
                      const ramcu_scoring_t *scoring, const ramcu_params_t *params,
                      uint8_t *master_right, int32_t *ext_len_right, int32_t *score_add_right,
                      ramcu_result_t *res_right, uint8_t *master_left, int32_t *ext_len_left,
                      int32_t *score_add_left, ramcu_result_t *res_left);

/* -outmat (ram_extend.c:1115-1132, bnw_extend.c:1027-1044): ramcu_extend() that also returns, for
 * every processed row r < res->rows_processed and every core i extendable in `direction`,
 *   path_state_out[(r * n_cores + i) * W + k]   state of the best path into band cell k (offset k - B) of
 *                                                the committed row: 0 substitution, 1 deletion, 2 insertion
 *                                                (the reference prints the base / '-' / the lowercase base)
 *   row_best_out[r * n_cores + i]               compute_nw_row()'s return value for the chosen base
 *   row_best_idx_out[r * n_cores + i]           band index of its first best cell
 *                                                (curr_row_best_col_idx - row_idx + BANDWIDTH)
 * with W = 2 * bandwidth + 1; the arrays hold n_rows rows (rows beyond are not reported: a caller that runs
 * ramcu_extend() first knows rows_processed and sizes them by that, as csrc/ramx_boundary.c does); entries of other
 * cores are left untouched. Debug path: n_rows x n_cores x W is limited to 16 Gi cells. */
int ramcu_extend_paths(ramcu_ctx *ctx, ramcu_library *lib, int direction, int n_cores,
                       const ramcu_core_t *cores, const ramcu_scoring_t *scoring,
                       const ramcu_params_t *params, uint8_t *master_out, int32_t *ext_len_out,
                       int32_t *score_add_out, ramcu_result_t *res, int n_rows, uint8_t *path_state_out,
                       int32_t *row_best_out, int32_t *row_best_idx_out);

/* VERBOSE >= 10 dumps (ram_extend.c:992-1090, 1139-1141, 1177-1214): ramcu_extend() that also returns, for the
 * first n_rows processed rows r, every core i extendable in `direction` and every candidate consensus base a,
 *   cand_out[((r * n_cores + i) * 4 + a) * 4 + 0]   compute_nw_row()'s return value (curr_row_best_score, unclamped)
 *                                            + 1   *max_score_sequence_idx (curr_row_best_col_idx)
 *                                            + 2   gap value of band cell 0      (score[row%2][n][0][1])
 *                                            + 3   gap value of band cell 2B     (score[row%2][n][2*BANDWIDTH][1])
 *   sums_out[r * 4 + a]                             score_given_cons
 *   cells_out (or NULL; VERBOSE >= 12)              [n_rows][n_cores][4][2B+1][2] = (sub, gap) of every band cell
 * everything a caller needs to print the reference's per-row dumps without any DP on the host
 * (csrc/ramx_boundary.c does). Entries of cores not extendable in `direction` are left untouched. */
int ramcu_extend_verbose(ramcu_ctx *ctx, ramcu_library *lib, int direction, int n_cores,
                         const ramcu_core_t *cores, const ramcu_scoring_t *scoring,
                         const ramcu_params_t *params, uint8_t *master_out, int32_t *ext_len_out,
                         int32_t *score_add_out, ramcu_result_t *res, int n_rows, int32_t *cand_out,
                         int32_t *sums_out, int32_t *cells_out);

/* Seed extension: the other callers of compute_nw_row() (bnw_extend.c:167-649), batched.
 * One ramcu_seed_t == one cons_seed_extend() call (bnw_extend.h:38-43): the query (host codes 0..3, any
 * other value is an N base) with its seed [query_seed_start, query_seed_end], the subject seed
 * [subj_seed_start, subj_seed_end] in library coordinates with the extent of the subject's sequence
 * (boundaries[] of the reference's seqLib) and its orientation. pair_seed_extend() (bnw_extend.h:31-36)
 * is the same call with the query taken from the library. Results are the reference's
 * left_max_score / right_max_score (their sum is the return value before complexity adjustment) and the
 * glocalSearchResult ranges. n_row[5] = score of an N query base against A,C,G,T,N
 * (scoreParams->matrix[SYM_N][..]). y_drop as in the reference: negative = no early stop. */
typedef struct
{
  const uint8_t *query;
  int32_t query_len, query_seed_start, query_seed_end;
  int32_t orient;
  int64_t subj_seed_start, subj_seed_end;
  int64_t lower_seq_bound, upper_seq_bound;
} ramcu_seed_t;
typedef struct
{
  int32_t left_score, right_score;
  int32_t query_start, query_end;
  int64_t subj_start, subj_end;
  int32_t left_rows, right_rows;   /* rows scanned before the y-drop break */
} ramcu_seed_result_t;
int ramcu_seed_extend_batch(ramcu_ctx *ctx, ramcu_library *lib, int n_seeds, const ramcu_seed_t *seeds,
                            const ramcu_scoring_t *scoring, const int32_t *n_row, int bandwidth, int y_drop,
                            ramcu_seed_result_t *results);
/* global_align() (bnw_extend.h:45-49): seq1 (host codes) is the given consensus, seq2 = library
 * positions [seq2_start, seq2_end]. */
int ramcu_global_align(ramcu_ctx *ctx, ramcu_library *lib, int orient, const uint8_t *seq1, int seq1_len,
                       int64_t seq2_start, int64_t seq2_end, const ramcu_scoring_t *scoring, const int32_t *n_row,
                       int bandwidth, int32_t *global_score, int32_t *local_score, int32_t *local_pos);

/* Test-only second boundary (compute_nw_row, bnw_extend.h:52-58). Runs the extension kernel for
 * `n_rows` rows with the consensus FORCED to forced_cons[r] (NULL = argmax as usual) and returns
 *   rows_out : [n_rows][n_cores][2*B+1][2] committed band rows ([..][0]=sub, [..][1]=gap);
 *              rows of instances not extendable in `direction` are left untouched
 *   best_out : [n_rows][n_cores][4] per-candidate best row score (before clamp/cap)
 *   sums_out : [n_rows][4] per-candidate column sums
 */
int ramcu_debug_rows(ramcu_ctx *ctx, ramcu_library *lib, int direction, int n_cores,
                     const ramcu_core_t *cores, const ramcu_scoring_t *scoring,
                     const ramcu_params_t *params, int n_rows, const uint8_t *forced_cons,
                     int32_t *rows_out, int32_t *best_out, int32_t *sums_out);

#ifdef __cplusplus
}
#endif
#endif
