/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY. Compiled ONLY into oracle/_ref/libramref.so.
 *
 * Flat-array entry points around the UNMODIFIED reference object code, so that tests and the
 * CPU-baseline leg of bench.py can call the reference's own extend_alignment()
 * (ram_extend.c:859) and compute_nw_row() (bnw_extend.c:750) without going through files.
 * It builds the reference's own structs (struct coreAlignment linked list, struct
 * sequenceLibrary, struct scoringSystem, int**** score) from flat inputs; all arithmetic is
 * done by the reference's objects. Includes the reference headers where they lie (-I$(REF)).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include "common.h"
#include "sequence.h"
#include "score_system.h"
#include "bnw_extend.h"
#include "ram_extend.h"

/* globals defined by ram_extend.c (ram_extend.c:40,52,61) */
extern int VERBOSE;
extern int WHEN_TO_STOP;
extern int l;

/* same layout as oracle_core_t / ramcu_core_t */
typedef struct
{
  int64_t left_seq_pos, right_seq_pos, lower_seq_bound, upper_seq_bound;
  int8_t orient, left_extendable, right_extendable, pad_[5];
} shim_core_t;

/* matrix[4][9] (columns 0..7, N) + gaps -> reference struct with a 100x100 int** matrix */
static struct scoringSystem *
make_scoring(const int32_t *m49, int gapopen, int gapextn)
{
  int i, j;
  struct scoringSystem *s = (struct scoringSystem *) calloc(1, sizeof(*s));
  s->matrix = (int **) malloc(100 * sizeof(int *));
  for (i = 0; i < 100; i++)
  {
    s->matrix[i] = (int *) malloc(100 * sizeof(int));
    for (j = 0; j < 100; j++)
      s->matrix[i][j] = -1;
  }
  for (i = 0; i < 4; i++)
  {
    for (j = 0; j < 8; j++)
      s->matrix[i][j] = m49[i * 9 + j];
    s->matrix[i][SYM_N] = m49[i * 9 + 8];
  }
  s->msize = 100;
  s->name = "shim";
  s->alphabet = "ACGTN";
  s->gapopen = gapopen;
  s->gapextn = gapextn;
  return s;
}

/* Export the reference's own tables: name in {14p43g,18p43g,20p43g,25p43g,repeatscout}. */
int
ramref_get_scoring(const char *name, int rs_match, int rs_mismatch, int rs_gap,
                   int32_t *m49, int32_t *gapopen, int32_t *gapextn)
{
  struct scoringSystem *s;
  int i, j;
  if (strcmp(name, "repeatscout") == 0)
    s = getRepeatScoutMatrix(rs_match, rs_mismatch, rs_gap);
  else if (!strcmp(name, "14p43g") || !strcmp(name, "18p43g") || !strcmp(name, "20p43g") ||
           !strcmp(name, "25p43g"))
    s = getMatrix((char *) name);
  else
    return -1;
  for (i = 0; i < 4; i++)
  {
    for (j = 0; j < 8; j++)
      m49[i * 9 + j] = s->matrix[i][j];
    m49[i * 9 + 8] = s->matrix[i][SYM_N];
  }
  *gapopen = s->gapopen;
  *gapextn = s->gapextn;
  freeScoringSystem(s);
  return 0;
}

static struct coreAlignment *
make_cores(int n, const shim_core_t *c)
{
  struct coreAlignment *arr = (struct coreAlignment *) calloc(n > 0 ? n : 1, sizeof(*arr));
  int i;
  for (i = 0; i < n; i++)
  {
    arr[i].next = (i + 1 < n) ? &arr[i + 1] : NULL;
    arr[i].seqIdx = i;
    arr[i].leftSeqPos = (uint64_t) c[i].left_seq_pos;
    arr[i].rightSeqPos = (uint64_t) c[i].right_seq_pos;
    arr[i].lowerSeqBound = (uint64_t) c[i].lower_seq_bound;
    arr[i].upperSeqBound = (uint64_t) c[i].upper_seq_bound;
    arr[i].leftExtendable = c[i].left_extendable;
    arr[i].rightExtendable = c[i].right_extendable;
    arr[i].orient = c[i].orient;
  }
  return arr;
}

/* The reference's extend_alignment() on flat inputs. master_out gets the chosen bases of rows
 * 0..L-1 in processing order (undone from the reference's master[] indexing, 1092-1095).
 * seconds (optional) = wall time of the extend_alignment() call alone. Returns extended bp. */
int
ramref_extend(int direction, int n, const shim_core_t *cores, const char *sequence,
              uint64_t seq_len, const int32_t *m49, int gapopen, int gapextn, int bandwidth,
              int cappenalty, int minimprovement, int L, int when_to_stop, int verbose,
              uint8_t *master_out, int32_t *ext_len_out, int32_t *score_add_out,
              double *seconds)
{
  struct coreAlignment *ca = make_cores(n, cores);
  struct scoringSystem *sc = make_scoring(m49, gapopen, gapextn);
  struct sequenceLibrary lib;
  int ****score = allocate_score(n, bandwidth);
  char *master = (char *) malloc((size_t) 2 * L + l + 2);
  struct timespec t0, t1;
  int bp, i;

  memset(&lib, 0, sizeof(lib));
  lib.sequence = (char *) sequence;
  lib.length = seq_len;
  lib.count = n;
  memset(master, 100, (size_t) 2 * L + l + 2);
  VERBOSE = verbose;
  WHEN_TO_STOP = when_to_stop;

  clock_gettime(CLOCK_MONOTONIC, &t0);
  bp = extend_alignment(direction, ca, score, &lib, master, bandwidth, cappenalty,
                        minimprovement, L, n, sc, NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds)
    *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  fflush(stdout);

  for (i = 0; i < L; i++)
    master_out[i] = (uint8_t) (direction ? master[L + l + i] : master[L - i - 1]);
  for (i = 0; i < n; i++)
  {
    ext_len_out[i] = direction ? ca[i].rightExtensionLen : ca[i].leftExtensionLen;
    score_add_out[i] = ca[i].score;
  }
  free_score(n, bandwidth, score);
  freeScoringSystem(sc);
  free(master);
  free(ca);
  return bp;
}

/* One call of the reference's compute_nw_row() on flat band rows (W = 2*bandwidth+1). */
int
ramref_nw_row(int direction, int row_idx, int cons_base, const shim_core_t *core,
              const int32_t *prev_sub, const int32_t *prev_gap, int32_t *cur_sub,
              int32_t *cur_gap, const char *sequence, const int32_t *m49, int gapopen,
              int gapextn, int bandwidth, int *best_idx)
{
  struct coreAlignment *ca = make_cores(1, core);
  struct scoringSystem *sc = make_scoring(m49, gapopen, gapextn);
  int ****score = allocate_score(1, bandwidth);
  int W = 2 * bandwidth + 1, k, best;
  int cur = row_idx % 2, prev = !cur;
  for (k = 0; k < W; k++)
  {
    score[prev][0][k][0] = prev_sub[k];
    score[prev][0][k][1] = prev_gap[k];
  }
  best = compute_nw_row(direction, row_idx, 0, (char) cons_base, ca, score,
                        ca->lowerSeqBound, ca->upperSeqBound, (char *) sequence, best_idx, sc,
                        bandwidth, 0, 0, NULL);
  for (k = 0; k < W; k++)
  {
    cur_sub[k] = score[cur][0][k][0];
    cur_gap[k] = score[cur][0][k][1];
  }
  free_score(1, bandwidth, score);
  freeScoringSystem(sc);
  free(ca);
  return best;
}

/* The reference loader on (2bit, ranges) -> flat arrays (sequence.c:505-923). Two-call
 * protocol: call with seq_out == NULL to get sizes, then with buffers. Returns n cores. */
int
ramref_load(const char *twobit, const char *ranges, int max_flanking_bp, int *n_out,
            uint64_t *seq_len_out, shim_core_t *cores_out, char *seq_out,
            uint64_t *offsets_out, uint64_t *boundaries_out)
{
  struct coreAlignment *ca = NULL, *c;
  int n = 0, i;
  struct sequenceLibrary *lib =
    loadSequenceSubsetMinimal((char *) twobit, (char *) ranges, &ca, &n, max_flanking_bp);
  *n_out = n;
  *seq_len_out = lib->length;
  if (seq_out)
  {
    memcpy(seq_out, lib->sequence, lib->length);
    for (c = ca, i = 0; c != NULL; c = c->next, i++)
    {
      memset(&cores_out[i], 0, sizeof(shim_core_t));
      cores_out[i].left_seq_pos = (int64_t) c->leftSeqPos;
      cores_out[i].right_seq_pos = (int64_t) c->rightSeqPos;
      cores_out[i].lower_seq_bound = (int64_t) c->lowerSeqBound;
      cores_out[i].upper_seq_bound = (int64_t) c->upperSeqBound;
      cores_out[i].orient = c->orient;
      cores_out[i].left_extendable = c->leftExtendable;
      cores_out[i].right_extendable = c->rightExtendable;
      if (offsets_out) offsets_out[i] = lib->offsets[i];
      if (boundaries_out) boundaries_out[i] = lib->boundaries[i];
    }
  }
  return n;
}

/* The reference's cons_seed_extend() (bnw_extend.c:196-484) on flat inputs: `name` selects one of the
 * reference's own scoring systems (lambda / background frequencies included, so the complexity
 * adjustment is the reference's too). out[7] = return value, query_start, query_end, subj_start,
 * subj_end, and -- recomputed with adjustment off -- the unadjusted score; n_row_out[5] = the N row. */
int
ramref_cons_seed_extend(const char *name, const char *query, int query_len, int query_seed_start,
                        int query_seed_end, const char *sequence, uint64_t seq_len, const uint64_t *boundaries,
                        uint64_t subj_seed_start, uint64_t subj_seed_end, int orient, int bandwidth,
                        int use_complexity_adjust, int y_drop, int64_t *out, int32_t *n_row_out)
{
  struct scoringSystem *s = getMatrix((char *) name);
  struct sequenceLibrary lib;
  struct glocalSearchResult res;
  int ****score = allocate_score(1, bandwidth);
  int nb = 0, sc, i;
  memset(&lib, 0, sizeof(lib));
  while (boundaries[nb] != 0)
    nb++;
  lib.sequence = (char *) sequence;
  lib.boundaries = (uint64_t *) boundaries;
  lib.count = nb;
  lib.length = seq_len;
  sc = cons_seed_extend((char *) query, query_len, query_seed_start, query_seed_end, &lib, subj_seed_start,
                        subj_seed_end, orient, score, s, bandwidth, &res, use_complexity_adjust, y_drop);
  out[0] = sc;
  out[1] = res.query_start;
  out[2] = res.query_end;
  out[3] = (int64_t) res.subj_start;
  out[4] = (int64_t) res.subj_end;
  out[5] = cons_seed_extend((char *) query, query_len, query_seed_start, query_seed_end, &lib, subj_seed_start,
                            subj_seed_end, orient, score, s, bandwidth, &res, 0, y_drop);
  for (i = 0; i < 4; i++)
    n_row_out[i] = s->matrix[SYM_N][i];
  n_row_out[4] = s->matrix[SYM_N][SYM_N];
  return sc;
}

/* The reference's global_align() (bnw_extend.c:486-649); it prints one line to stdout. */
int
ramref_global_align(const char *name, int orient, uint64_t seq1_start, uint64_t seq1_end, uint64_t seq2_start,
                    uint64_t seq2_end, const char *sequence, uint64_t seq_len, int bandwidth)
{
  struct scoringSystem *s = getMatrix((char *) name);
  struct sequenceLibrary lib;
  uint64_t boundaries[2];
  int ****score = allocate_score(1, bandwidth);
  memset(&lib, 0, sizeof(lib));
  boundaries[0] = seq_len;
  boundaries[1] = 0;
  lib.sequence = (char *) sequence;
  lib.boundaries = boundaries;
  lib.count = 1;
  lib.length = seq_len;
  return global_align(orient, seq1_start, seq1_end, seq1_start, seq2_start, seq2_end, seq2_start, score, 8, &lib, s,
                      bandwidth);
}
