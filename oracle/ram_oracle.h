/*
 * ram_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of RAMExtend's banded-DP extension path (RepeatAfterMe 0.0.7):
 *   compute_nw_row()      /root/reference/bnw_extend.c:750-1048
 *   extend_alignment()    /root/reference/ram_extend.c:859-1258
 *   scoring tables        /root/reference/score_system.c:91-172, 182-400
 *   base complement       /root/reference/sequence.c:1140-1160
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product path (repeatafterme_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py checks this restatement against the
 * reference's own object code (oracle/_ref/libramref.so, built from /root/reference by
 * oracle/Makefile) row by row and end to end, against the reference's unit-test known-answer
 * vectors (bnw_extend.c:1566-1567) and against committed golden fixtures (tests/golden/).
 */
#ifndef RAM_ORACLE_H
#define RAM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_NEG      (-1000000000)   /* bnw_extend.c:760,802-804 */
#define ORACLE_OOB      (-987654321)    /* bnw_extend.c:767        */
#define ORACLE_NCODES   9               /* columns 0..7 (ACGTacgt) + N(99) folded to 8 */

/* One MSA instance == one struct coreAlignment (common.h:80-98), flattened. */
typedef struct
{
  int64_t left_seq_pos;      /* leftSeqPos  : index into the library's sequence[] */
  int64_t right_seq_pos;     /* rightSeqPos */
  int64_t lower_seq_bound;   /* lowerSeqBound (inclusive) */
  int64_t upper_seq_bound;   /* upperSeqBound (inclusive) */
  int8_t orient;             /* 1 = reverse strand */
  int8_t left_extendable;
  int8_t right_extendable;
  int8_t pad_[5];
} oracle_core_t;

/* struct scoringSystem (score_system.h:7-17) reduced to what the path reads:
 * matrix[cons 0..3][seq code], seq code = 0..7 as in sequence.h:7-15, 8 = SYM_N (99). */
typedef struct
{
  int32_t matrix[4][ORACLE_NCODES];
  int32_t gapopen;
  int32_t gapextn;
} oracle_scoring_t;

typedef struct
{
  int32_t extended_bp;     /* return value of extend_alignment(): max_row + 1          */
  int32_t rows_processed;  /* iterations of the column loop executed                    */
  int32_t stop_row;        /* row_idx at the WHEN_TO_STOP break, or -1 if L ran out     */
  int32_t hit_limit;       /* 1 if the reference prints the "to the limit" WARNING      */
  int32_t max_score;       /* max_extension_score at exit                               */
  int32_t n_extendable;    /* instances extendable in this direction                    */
} oracle_result_t;

/* Fill a scoring system by name: "14p43g","18p43g","20p43g","25p43g" (score_system.c:207-376)
 * or "repeatscout" (match/mismatch/gap; score_system.c:91-172). Returns 0, or -1 if unknown. */
int oracle_get_scoring(const char *name, int rs_match, int rs_mismatch, int rs_gap,
                       oracle_scoring_t *out);

/* sequence byte (0..7, 99, anything) -> matrix column 0..8 */
int oracle_seq_code(int byte_value);
/* compl() of sequence.c:1140-1160 on matrix-column codes */
int oracle_compl_code(int code);

/* compute_nw_row(): one band row for one instance under consensus hypothesis cons_base.
 * prev_sub/prev_gap and cur_sub/cur_gap are W=2*bandwidth+1 ints, index k = offset+bandwidth.
 * Returns best row score; *best_idx = row_idx + offset of the first maximal cell. */
int oracle_nw_row(int direction, int row_idx, int cons_base, const oracle_core_t *core,
                  const int32_t *prev_sub, const int32_t *prev_gap,
                  int32_t *cur_sub, int32_t *cur_gap,
                  const uint8_t *sequence, const oracle_scoring_t *sc, int bandwidth,
                  int *best_idx);

/* extend_alignment(): whole column loop for one direction.
 *   master_out   : >= L bytes; master_out[r] = chosen base (0..3) of processed row r
 *   ext_len_out  : n ints; new left/rightExtensionLen, or 0 when the reference leaves it alone
 *   score_add_out: n ints; amount added to coreAlignment.score
 *   trace_out    : optional (may be NULL), 6 ints per processed row:
 *                  sum[A],sum[C],sum[G],sum[T],besta,max_row_after
 * Returns 0. */
int oracle_extend(int direction, int n, const oracle_core_t *cores,
                  const uint8_t *sequence, const oracle_scoring_t *sc,
                  int bandwidth, int cappenalty, int minimprovement, int L, int when_to_stop,
                  uint8_t *master_out, int32_t *ext_len_out, int32_t *score_add_out,
                  oracle_result_t *res, int32_t *trace_out);

/* cell updates executed for a result: rows_processed * n_extendable * (2B+1) * 4 */
double oracle_cell_updates(const oracle_result_t *res, int bandwidth);

#ifdef __cplusplus
}
#endif
#endif
