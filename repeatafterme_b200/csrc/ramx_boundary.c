/*
 * ramx_boundary.c -- host-side C boundary: a strong definition of the reference's
 *
 *     int extend_alignment(int direction, struct coreAlignment *coreAlign, int ****score,
 *                          struct sequenceLibrary *seqLib, char *master, int BANDWIDTH,
 *                          int CAPPENALTY, int MINIMPROVEMENT, int L, int N,
 *                          struct scoringSystem *scoreParams, FILE *pathStringFile);
 *
 * (ram_extend.h:9-13, defined in ram_extend.c:859-1258). The product CLI is linked from the
 * reference's own, unmodified host objects (compiled out of tree from /root/reference) with this
 * one symbol weakened in ram_extend.o (objcopy --weaken-symbol=extend_alignment), so that the
 * reference's main() -- flag parsing, 2bit loader, overlap avoidance, report, every output --
 * runs unchanged and only the column loop goes to the GPU through include/ramcu.h.
 *
 * What stays on the host here: walking the linked list (O(N)), folding the 100x100 int** matrix
 * to the 4x5 table the path reads, the VERBOSE<=3 prints of ram_extend.c:886-892,1218-1231, and
 * the write-back of master[] / per-core fields (ram_extend.c:1092-1095,1234-1247). No DP.
 * There is no CPU fallback: a CUDA failure prints the error and exits, in the reference's style.
 *
 * Pair launch: main() calls right, narrows the bounds of cores a right extension ran into
 * (ram_extend.c:445-499), then calls left. The right call here runs BOTH passes in one launch
 * (ramcu_extend_pair: the passes hide each other's per-column exchange) with the bounds as they
 * are then, and keeps the left results together with the flattened cores they were computed
 * from. The left call flattens the cores again; if they are byte-identical (no bound was
 * narrowed) and nothing else changed, the kept results ARE the left pass; otherwise the left
 * pass is run again on its own. RAMCU_PAIR=0 disables this (two launches, as before).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <ctype.h>
#include <time.h>

#include "common.h"        /* struct coreAlignment        (reference, -I$(REF)) */
#include "sequence.h"      /* struct sequenceLibrary, SYM_N */
#include "score_system.h"  /* struct scoringSystem */
#include "ram_extend.h"

#include "ramcu.h"
#include "ramx_loader.h"
#include "ramx_boundary.h"

/* globals of ram_extend.c that extend_alignment() reads (ram_extend.c:40,52,61) */
extern int VERBOSE;
extern int WHEN_TO_STOP;
extern int l;

static ramcu_ctx *g_ctx = NULL;
static ramcu_library *g_lib = NULL;
static const char *g_lib_seq = NULL;
static uint64_t g_lib_len = 0;

/* left-pass results computed ahead by the right call, and what they were computed from */
static struct
{
  int valid, n, L;
  ramcu_core_t *cores;
  ramcu_scoring_t sc;
  ramcu_params_t pr;
  const ramcu_library *lib;
  uint8_t *cons;
  int32_t *ext_len, *score_add;
  ramcu_result_t res;
} g_ahead;

static double
now_s(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

static void
ahead_drop(void)
{
  free(g_ahead.cores);
  free(g_ahead.cons);
  free(g_ahead.ext_len);
  free(g_ahead.score_add);
  memset(&g_ahead, 0, sizeof(g_ahead));
}

/* Batch front-end (csrc/ramx_batch.c): when set, every extend_alignment() call that is not an -outmat
 * call is offered to the hook first, with the flattened inputs; it returns 1 after filling the outputs
 * (or does not return: the recording phase leaves main() with longjmp), 0 to let the call run here. */
ramx_extend_hook_fn ramx_extend_hook = NULL;

static void
die(const char *what)
{
  printf("RAMExtend: CUDA extension failed in %s: %s\n", what, ramcu_last_error());
  exit(1);
}

/* the right and the left pass share one library upload */
static void
ensure_library(struct sequenceLibrary *seqLib)
{
  const int timing = getenv("RAMX_TIMING") != NULL;
  const double t0 = now_s();
  if (g_lib == NULL || g_lib_seq != seqLib->sequence || g_lib_len != seqLib->length)
  {
    ahead_drop();
    if (g_lib)
      ramcu_library_free(g_ctx, g_lib);
    g_lib = NULL;
    {
      /* a library built by the product loader travels as the .2bit file's own packed bytes and is
       * decoded on the device; any other library (1 byte per base) is packed on the device */
      const struct ramx_twobit_source *src =
        getenv("RAMCU_BYTE_LIBRARY") ? NULL : ramx_loader_source(seqLib->sequence, seqLib->length);
      if (src != NULL)
      {
        if (ramcu_library_from_twobit(g_ctx, src->blob, src->blob_bytes, src->n_windows,
                                      (const ramcu_window_t *) src->windows, src->n_nranges, src->n_start,
                                      src->n_len, src->total, &g_lib) != RAMCU_OK)
          die("ramcu_library_from_twobit");
      }
      else if (ramcu_library_upload(g_ctx, (const uint8_t *) seqLib->sequence, seqLib->length, &g_lib) != RAMCU_OK)
        die("ramcu_library_upload");
    }
    g_lib_seq = seqLib->sequence;
    g_lib_len = seqLib->length;
    if (timing)
      fprintf(stderr, "[ramx timing] library to the device %.3f s\n", now_s() - t0);
  }
}

/* forget the library of the previous main() run (its host buffer may have been freed and the address reused) */
void
ramx_boundary_forget_library(void)
{
  ahead_drop();
  if (g_lib && g_ctx)
    ramcu_library_free(g_ctx, g_lib);
  g_lib = NULL;
  g_lib_seq = NULL;
  g_lib_len = 0;
}

/* RAMX_TIMING: where the wall clock of the process goes outside the calls timed above. One handler is registered
 * when the library is loaded (it runs last, after the CUDA runtime's own exit handler), one after the context
 * exists (it runs first): the difference is the teardown of the context. */
static double g_t_start;

static void
timing_at_exit_first(void)
{
  fflush(stdout);
  fprintf(stderr, "[ramx timing] main() done %.3f s after process start\n", now_s() - g_t_start);
}

static void
timing_at_exit_last(void)
{
  fprintf(stderr, "[ramx timing] exit handlers done %.3f s after process start\n", now_s() - g_t_start);
}

__attribute__((constructor)) static void
timing_init(void)
{
  g_t_start = now_s();
  if (getenv("RAMX_TIMING") != NULL)
    atexit(timing_at_exit_last);
}

ramcu_ctx *
ramx_boundary_context(void)
{
  if (g_ctx == NULL)
  {
    const char *dev = getenv("RAMCU_DEVICE");
    if (ramcu_init(dev ? atoi(dev) : 0, &g_ctx) != RAMCU_OK)
      die("ramcu_init");
    if (getenv("RAMX_TIMING") != NULL)
      atexit(timing_at_exit_first);
  }
  return g_ctx;
}

/* reference host function (report.c:19-20), linked unchanged */
void printExtensionRegion(struct coreAlignment *currCore, int direction, int row, int capped,
                          struct sequenceLibrary *seqLib);

/*
 * -vvvv / -vvvvv / -vvvvvv (VERBOSE 10 / 12 / 20): the per-row dumps of the reference's column loop
 * (ram_extend.c:949-959, 992-1090, 1177-1214; bnw_extend.c:843-868), printed in the reference's order
 * from what the kernel reports for every candidate base of every processed row (ramcu_extend_verbose):
 * row best and its column, the gap values of the two edge cells, the column sums and -- from -vvvvv --
 * every band cell. The only arithmetic done here is the bookkeeping the prints themselves show
 * (clamp / cap of a printed score, the record setter, the improvement rule); no DP.
 */
static void
print_verbose_rows(int direction, struct coreAlignment *coreAlign, struct sequenceLibrary *seqLib, int n,
                   int BANDWIDTH, int CAPPENALTY, int MINIMPROVEMENT, int rows, const uint8_t *cons,
                   const int32_t *cand, const int32_t *sums, const int32_t *cells)
{
  const size_t W = 2 * (size_t) BANDWIDTH + 1;
  int *high = (int *) calloc(n > 0 ? n : 1, sizeof(int));
  int max_ext = 0, max_row = -1;
  int r, i, offset;
  struct coreAlignment *c;
  if (!high)
  {
    printf("extend_alignment: Error could not allocate memory for the verbose dumps\n");
    exit(1);
  }
  for (r = 0; r < rows; r++)
  {
    int curr = 0, a;
    char besta = 0;
    int totalExtendable = 0, numOutOfSeq = 0, numHighScore = 0;
    float score_per_position_since_max = 0.0;
    for (a = 0; a < 4; a++)
    {
      i = 0;
      for (c = coreAlign; c != NULL && i < n; c = c->next, i++)
      {
        const int32_t *cd = cand + (((size_t) r * n + i) * 4 + a) * 4;
        int best = cd[0];
        if (!((direction && c->rightExtendable) || (!direction && c->leftExtendable)))
          continue;
        if (direction)
          printf("RIGHT ROW %d with '%c': n = %d\n", r, num_to_char((char) a), i);
        else
          printf("LEFT ROW %d with '%c': n = %d\n", r, num_to_char((char) a), i);
        if (VERBOSE > 12 && VERBOSE < 100)
        {
          /* the bound tests of compute_nw_row(), bnw_extend.c:778-788, 824-868, for their messages */
          uint64_t start_seq_pos;
          if (direction)
            start_seq_pos = c->orient ? c->rightSeqPos - 1 : c->rightSeqPos + 1;
          else
            start_seq_pos = c->orient ? c->leftSeqPos + 1 : c->leftSeqPos - 1;
          for (offset = -BANDWIDTH; offset <= BANDWIDTH; offset++)
          {
            const int seq_offset = ((direction ? 1 : 0) == (c->orient ? 1 : 0)) ? -offset - r : offset + r;
            const uint64_t seq_idx = start_seq_pos + seq_offset;
            if (start_seq_pos < (uint64_t) BANDWIDTH && seq_offset < 0 && (uint64_t) abs(seq_offset) > start_seq_pos)
              printf("Short circuiting right extension n=%d ( start_seq_pos = %ld, "
                     "seq_offset = %d ) because of lower sequence bound %ld [overflow avoided]\n", i,
                     (long) start_seq_pos, seq_offset, (long) c->lowerSeqBound);
            else if (seq_idx > c->upperSeqBound)
              printf("Short circuiting right extension n=%d ( start_seq_pos = %ld, "
                     "seq_offset = %d, seq_idx = %ld ) because of upper sequence bound %ld\n", i,
                     (long) start_seq_pos, seq_offset, (long) seq_idx, (long) c->upperSeqBound);
            else if (seq_idx < c->lowerSeqBound)
              printf("Short circuiting right extension n=%d ( start_seq_pos = %ld, "
                     "seq_offset = %d, seq_idx = %ld ) because of lower sequence bound %ld\n", i,
                     (long) start_seq_pos, seq_offset, (long) seq_idx, (long) c->lowerSeqBound);
          }
        }
        if (VERBOSE >= 12 && cells != NULL)
        {
          const int32_t *cl = cells + (((size_t) r * n + i) * 4 + a) * W * 2;
          printf("    Gap: ");
          for (offset = 0; offset < (int) W; offset++)
            printf(" %d", cl[2 * offset + 1]);
          printf("\n");
          printf("    Sub: ");
          for (offset = 0; offset < (int) W; offset++)
            printf(" %d", cl[2 * offset]);
          printf("\n");
        }
        if (best < 0)
          printf("  best score = %d @ column %d -- max(0,best_score) = 0!, prev best score = %d\n", best, cd[1],
                 high[i]);
        else
          printf("  best score = %d @ column %d, prev best score = %d\n", best, cd[1], high[i]);
        if ((direction && cd[2] < -279000) || (!direction && cd[3] < -279000))
          printf(" **OUT_OF_SEQ**");
        if (best < 0)
          best = 0;
        if (!(best >= high[i] + CAPPENALTY))
          printf(" **CAPPED** contributing = %d", high[i] + CAPPENALTY);
        printf("\n");
        if (VERBOSE >= 12)
          printExtensionRegion(c, direction, r, 0, seqLib);
      }
      printf("  Total Score for '%c' = %d\n", num_to_char((char) a), sums[(size_t) r * 4 + a]);
      if (sums[(size_t) r * 4 + a] > curr)
      {
        curr = sums[(size_t) r * 4 + a];
        besta = (char) a;
      }
    }
    if ((uint8_t) besta != cons[r])
    {
      printf("extend_alignment: the verbose dump and the extension disagree on row %d\n", r);
      exit(1);
    }
    printf("ROW %d complete, '%c' chosen as the consensus. curr_ext_score = %d\n", r, num_to_char(besta), curr);
    /* the bookkeeping of the reference's fifth pass (ram_extend.c:1100-1150) */
    i = 0;
    for (c = coreAlign; c != NULL && i < n; c = c->next, i++)
    {
      const int32_t *cd = cand + (((size_t) r * n + i) * 4 + besta) * 4;
      if (!((direction && c->rightExtendable) || (!direction && c->leftExtendable)))
        continue;
      if ((!direction && cd[2] < -279000) || (direction && cd[3] < -279000))
        numOutOfSeq++;
      if (cd[0] > high[i])
      {
        high[i] = cd[0];
        numHighScore++;
      }
      totalExtendable++;
    }
    if (abs(r - max_row) > 0)
      score_per_position_since_max = ((float) (curr - max_ext) / abs(r - max_row));
    printf("Alignment Extension: curr_extension_score = %d, "
           "max_extension_score = %d @ row %d\n                     "
           "score_per_position_overall = %0.1f, "
           "score_per_position_since_max = %0.1f\n"
           "                     total_edges = %d, num_extending = %d, num_out_of_seq = %d\n",
           curr, max_ext, max_row, ((float) curr / (r + 1)), score_per_position_since_max, totalExtendable,
           numHighScore, numOutOfSeq);
    if (curr >= max_ext + (abs(max_row - r) * MINIMPROVEMENT))
    {
      printf("                     **This row is now the new max**\n");
      max_row = r;
      max_ext = curr;
    }
    else
      printf("                     extensions since last max score %d\n", abs(r - max_row));
  }
  free(high);
}

/* one extend_alignment() call with the -vvvv.. dumps */
static void
run_verbose(int direction, struct coreAlignment *coreAlign, struct sequenceLibrary *seqLib, int n, int N,
            int BANDWIDTH, int CAPPENALTY, int MINIMPROVEMENT, struct scoringSystem *scoreParams,
            const ramcu_core_t *cores, const ramcu_scoring_t *scp, const ramcu_params_t *prp, uint8_t *cons,
            int32_t *ext_len, int32_t *score_add, ramcu_result_t *resp)
{
  /* -vvvv and up: a plain run tells how many rows there are, then the run whose results are used also
   * returns the dump data of exactly those rows */
  int32_t *cand, *sums, *cells = NULL;
  size_t rows;
  ahead_drop();
  if (VERBOSE >= 12)
  {
    /* ram_extend.c:949-959: the boundary row, for all N cores */
    int i, offset;
    for (i = 0; i < N; i++)
    {
      printf("SW Matrix Boundary Conditions ( n = %d ):\n", i);
      printf("  GAP: ");
      for (offset = -BANDWIDTH; offset <= BANDWIDTH; offset++)
        printf(" %d", offset == 0 ? 0 : abs(offset) * scoreParams->gapextn + scoreParams->gapopen);
      printf("\n  SUB: ");
      for (offset = -BANDWIDTH; offset <= BANDWIDTH; offset++)
        printf(" %d", offset == 0 ? 0 : abs(offset) * scoreParams->gapextn + scoreParams->gapopen);
      printf("\n");
    }
  }
  if (ramcu_extend(g_ctx, g_lib, direction ? 1 : 0, n, cores, scp, prp, cons, ext_len, score_add, resp) != RAMCU_OK)
    die("ramcu_extend");
  rows = (size_t) (resp->rows_processed > 0 ? resp->rows_processed : 1);
  cand = (int32_t *) calloc(rows * (size_t) (n > 0 ? n : 1) * 16, sizeof(int32_t));
  sums = (int32_t *) calloc(rows * 4, sizeof(int32_t));
  if (VERBOSE >= 12)
    cells = (int32_t *) calloc(rows * (size_t) (n > 0 ? n : 1) * 4 * (2 * (size_t) BANDWIDTH + 1) * 2, sizeof(int32_t));
  if (!cand || !sums || (VERBOSE >= 12 && !cells))
  {
    printf("extend_alignment: Error could not allocate memory for the verbose dumps\n");
    exit(1);
  }
  if (ramcu_extend_verbose(g_ctx, g_lib, direction ? 1 : 0, n, cores, scp, prp, cons, ext_len, score_add, resp,
                           resp->rows_processed, cand, sums, cells) != RAMCU_OK)
    die("ramcu_extend_verbose");
  print_verbose_rows(direction, coreAlign, seqLib, n, BANDWIDTH, CAPPENALTY, MINIMPROVEMENT, resp->rows_processed, cons,
                     cand, sums, cells);
  free(cand);
  free(sums);
  free(cells);
}

int
extend_alignment(int direction, struct coreAlignment *coreAlign, int ****score,
                 struct sequenceLibrary *seqLib, char *master, int BANDWIDTH, int CAPPENALTY,
                 int MINIMPROVEMENT, int L, int N, struct scoringSystem *scoreParams,
                 FILE *pathStringFile)
{
  struct coreAlignment *c;
  ramcu_core_t *cores;
  ramcu_scoring_t sc;
  ramcu_params_t pr;
  ramcu_result_t res;
  uint8_t *cons;
  int32_t *ext_len, *score_add;
  int n, a, b, r;
  (void) score; /* the int**** scratch of allocate_score() is not used by the GPU path */

  if (VERBOSE >= 3)
  {
    if (direction)
      printf("extend_alignment(right): Called with %d edges\n", N);
    else
      printf("extend_alignment(left): Called with %d edges\n", N);
  }
  const int timing = getenv("RAMX_TIMING") != NULL;
  double t0 = now_s();
  int handled = 0;
  if (g_ctx == NULL)
  {
    ramx_boundary_context();
    if (timing)
      fprintf(stderr, "[ramx timing] CUDA context %.3f s\n", now_s() - t0);
    t0 = now_s();
  }
  /* matrix[cons][A,C,G,T,N]; the soft-masked columns 4..7 must score like N (they do for every
   * matrix score_system.c builds), because the device library keeps 2 bits + an N flag per base */
  memset(&sc, 0, sizeof(sc));
  for (a = 0; a < 4; a++)
  {
    for (b = 0; b < 4; b++)
      sc.matrix[a][b] = scoreParams->matrix[a][b];
    sc.matrix[a][4] = scoreParams->matrix[a][SYM_N];
  }
  sc.gapopen = scoreParams->gapopen;
  sc.gapextn = scoreParams->gapextn;

  n = 0;
  for (c = coreAlign; c != NULL && n < N; c = c->next)
    n++;
  cores = (ramcu_core_t *) calloc(n > 0 ? n : 1, sizeof(ramcu_core_t));
  cons = (uint8_t *) malloc(L > 0 ? L : 1);
  ext_len = (int32_t *) calloc(n > 0 ? n : 1, sizeof(int32_t));
  score_add = (int32_t *) calloc(n > 0 ? n : 1, sizeof(int32_t));
  if (!cores || !cons || !ext_len || !score_add)
  {
    printf("extend_alignment: Error could not allocate memory for the flattened cores\n");
    exit(1);
  }
  n = 0;
  for (c = coreAlign; c != NULL && n < N; c = c->next, n++)
  {
    cores[n].left_seq_pos = (int64_t) c->leftSeqPos;
    cores[n].right_seq_pos = (int64_t) c->rightSeqPos;
    cores[n].lower_seq_bound = (int64_t) c->lowerSeqBound;
    cores[n].upper_seq_bound = (int64_t) c->upperSeqBound;
    cores[n].orient = c->orient ? 1 : 0;
    cores[n].left_extendable = c->leftExtendable ? 1 : 0;
    cores[n].right_extendable = c->rightExtendable ? 1 : 0;
  }

  memset(&pr, 0, sizeof(pr));
  pr.bandwidth = BANDWIDTH;
  pr.cappenalty = CAPPENALTY;
  pr.minimprovement = MINIMPROVEMENT;
  pr.L = L;
  pr.when_to_stop = WHEN_TO_STOP;

  if (ramx_extend_hook != NULL && pathStringFile == NULL && VERBOSE < 10)
    handled = ramx_extend_hook(direction ? 1 : 0, n, cores, &sc, &pr, seqLib, cons, ext_len, score_add, &res);
  if (!handled)
    ensure_library(seqLib);
  t0 = now_s();

  if (handled)
  {
    /* answered by the batch front-end */
  }
  else if (pathStringFile != NULL)
  {
    /* -outmat: the kernel also reports the path state of every band cell of every committed row;
     * the lines of ram_extend.c:1122-1132 are written here, in the reference's order */
    const size_t W = 2 * (size_t) BANDWIDTH + 1;
    size_t cells;
    uint8_t *state;
    int32_t *rbest, *ridx;
    char *path = (char *) malloc(W + 1);
    ahead_drop();
    /* a plain run tells how many rows there are; the path state is then sized by that, not by L */
    if (ramcu_extend(g_ctx, g_lib, direction ? 1 : 0, n, cores, &sc, &pr, cons, ext_len, score_add, &res) != RAMCU_OK)
      die("ramcu_extend");
    cells = (size_t) (res.rows_processed > 0 ? res.rows_processed : 1) * (size_t) (n > 0 ? n : 1);
    state = (uint8_t *) calloc(cells * W, 1);
    rbest = (int32_t *) calloc(cells, sizeof(int32_t));
    ridx = (int32_t *) calloc(cells, sizeof(int32_t));
    if (!state || !rbest || !ridx || !path)
    {
      printf("extend_alignment: Error could not allocate memory for the path strings\n");
      exit(1);
    }
    if (ramcu_extend_paths(g_ctx, g_lib, direction ? 1 : 0, n, cores, &sc, &pr, cons, ext_len, score_add,
                           &res, res.rows_processed, state, rbest, ridx) != RAMCU_OK)
      die("ramcu_extend_paths");
    for (r = 0; r < res.rows_processed; r++)
    {
      int i = 0;
      for (c = coreAlign; c != NULL && i < n; c = c->next, i++)
      {
        uint64_t start_seq_pos;
        int offset;
        if (!((direction && c->rightExtendable) || (!direction && c->leftExtendable)))
          continue;
        /* bnw_extend.c:778-788, 824-837 */
        if (direction)
          start_seq_pos = c->orient ? c->rightSeqPos - 1 : c->rightSeqPos + 1;
        else
          start_seq_pos = c->orient ? c->leftSeqPos + 1 : c->leftSeqPos - 1;
        for (offset = -BANDWIDTH; offset < BANDWIDTH; offset++)   /* the last cell is cut off (:1124) */
        {
          const int seq_offset = ((direction ? 1 : 0) == (c->orient ? 1 : 0)) ? -offset - r : offset + r;
          const uint64_t seq_idx = start_seq_pos + seq_offset;
          const uint8_t st = state[((size_t) r * n + i) * W + (size_t) (offset + BANDWIDTH)];
          char base = 'X';
          if (seq_idx >= c->lowerSeqBound && seq_idx <= c->upperSeqBound)
            base = num_to_char(seqLib->sequence[seq_idx]);
          path[offset + BANDWIDTH] = (base == 'X' || st == 0) ? base : (st == 1) ? '-' : (char) tolower(base);
        }
        path[2 * BANDWIDTH] = '\0';
        fprintf(pathStringFile, direction ? "dir=right" : "dir=left");
        fprintf(pathStringFile, ":n=%d:row=%d: %s best:score=%d:offset=%d:cons=%c\n", i, r, path,
                rbest[(size_t) r * n + i], ridx[(size_t) r * n + i], num_to_char((char) cons[r]));
      }
    }
    free(state);
    free(rbest);
    free(ridx);
    free(path);
    if (VERBOSE >= 10)   /* -outmat together with -vvvv..: the dumps go to stdout, same results */
      run_verbose(direction, coreAlign, seqLib, n, N, BANDWIDTH, CAPPENALTY, MINIMPROVEMENT, scoreParams, cores, &sc, &pr,
                  cons, ext_len, score_add, &res);
  }
  else if (VERBOSE >= 10)
    run_verbose(direction, coreAlign, seqLib, n, N, BANDWIDTH, CAPPENALTY, MINIMPROVEMENT, scoreParams, cores, &sc, &pr,
                cons, ext_len, score_add, &res);
  else
  {
    const char *pe = getenv("RAMCU_PAIR");
    const int pair_ok = !(pe && atoi(pe) == 0);
    if (direction && pair_ok)
    {
      /* both passes now; the left results wait for the left call */
      ahead_drop();
      g_ahead.cores = (ramcu_core_t *) malloc((n > 0 ? n : 1) * sizeof(ramcu_core_t));
      g_ahead.cons = (uint8_t *) malloc(L > 0 ? L : 1);
      g_ahead.ext_len = (int32_t *) calloc(n > 0 ? n : 1, sizeof(int32_t));
      g_ahead.score_add = (int32_t *) calloc(n > 0 ? n : 1, sizeof(int32_t));
      if (!g_ahead.cores || !g_ahead.cons || !g_ahead.ext_len || !g_ahead.score_add)
      {
        printf("extend_alignment: Error could not allocate memory for the left-pass results\n");
        exit(1);
      }
      if (ramcu_extend_pair(g_ctx, g_lib, n, cores, &sc, &pr, cons, ext_len, score_add, &res,
                            g_ahead.cons, g_ahead.ext_len, g_ahead.score_add, &g_ahead.res) != RAMCU_OK)
        die("ramcu_extend_pair");
      memcpy(g_ahead.cores, cores, (size_t) n * sizeof(ramcu_core_t));
      g_ahead.n = n;
      g_ahead.L = L;
      g_ahead.sc = sc;
      g_ahead.pr = pr;
      g_ahead.lib = g_lib;
      g_ahead.valid = 1;
    }
    else if (!direction && g_ahead.valid && g_ahead.n == n && g_ahead.L == L && g_ahead.lib == g_lib &&
             memcmp(&g_ahead.sc, &sc, sizeof(sc)) == 0 && memcmp(&g_ahead.pr, &pr, sizeof(pr)) == 0 &&
             memcmp(g_ahead.cores, cores, (size_t) n * sizeof(ramcu_core_t)) == 0)
    {
      /* overlap avoidance changed nothing: the left pass has been computed already */
      memcpy(cons, g_ahead.cons, L > 0 ? L : 1);
      memcpy(ext_len, g_ahead.ext_len, (size_t) n * sizeof(int32_t));
      memcpy(score_add, g_ahead.score_add, (size_t) n * sizeof(int32_t));
      res = g_ahead.res;
      ahead_drop();
    }
    else
    {
      ahead_drop();
      if (ramcu_extend(g_ctx, g_lib, direction ? 1 : 0, n, cores, &sc, &pr, cons, ext_len, score_add,
                       &res) != RAMCU_OK)
        die("ramcu_extend");
    }
  }

  if (timing)
    fprintf(stderr, "[ramx timing] extend_alignment(%s) %.3f s (kernel %.1f ms)\n", direction ? "right" : "left",
            now_s() - t0, res.kernel_ms);
  /* consensus of every processed row (ram_extend.c:1092-1095) */
  for (r = 0; r < res.rows_processed; r++)
  {
    if (direction)
      master[L + l + r] = (char) cons[r];
    else
      master[L - r - 1] = (char) cons[r];
  }
  if (res.stop_row >= 0 && VERBOSE >= 3)
    printf("Ending...due to row_idx=%d - max_extension_score_row_idx=%d <= -WHEN_TO_STOP=%d\n",
           res.stop_row, res.extended_bp - 1, WHEN_TO_STOP);
  if (res.hit_limit)
  {
    if (direction)
      printf("WARNING: Extended sequence right to the limit ( L=%d ).\n", L);
    else
      printf("WARNING: Extended sequence left to the limit ( L=%d ).\n", L);
  }
  /* per-core extension length and score (ram_extend.c:1234-1247) */
  n = 0;
  for (c = coreAlign; c != NULL && n < N; c = c->next, n++)
  {
    if (ext_len[n] > 0)
    {
      if (direction)
        c->rightExtensionLen = ext_len[n];
      else
        c->leftExtensionLen = ext_len[n];
      c->score += score_add[n];
    }
  }
  free(cores);
  free(cons);
  free(ext_len);
  free(score_add);
  return res.extended_bp;
}
