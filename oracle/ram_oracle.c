/*
 * ram_oracle.c -- TEST INFRASTRUCTURE ONLY (see ram_oracle.h).
 *
 * Scalar, single-threaded CPU restatement of the RAMExtend banded-DP extension path, written
 * against flat arrays instead of the reference's linked list / int**** state. It keeps the
 * reference's evaluation order (ascending offsets, candidates A,C,G,T, then a commit pass) so
 * that every intermediate it exposes can be diffed against the reference's object code.
 *
 * Parity status: PINNED (see header).
 */
#include <stdlib.h>
#include <string.h>
#include "ram_oracle.h"

/* ------------------------------------------------------------------ scoring tables ------- */

/* matrix[consensus][sequence] for A,C,G,T; facts restated from score_system.c:207-376.
 * NB the tables are not symmetric (e.g. 14p43g A->G -10 but G->A -7). */
static const int8_t TAB_14p43g[4][4] = {{9, -18, -10, -21}, {-18, 11, -18, -7}, {-7, -18, 11, -18}, {-21, -10, -18, 9}};
static const int8_t TAB_18p43g[4][4] = {{9, -15, -8, -18}, {-16, 10, -16, -5}, {-5, -16, 10, -16}, {-18, -8, -15, 9}};
static const int8_t TAB_20p43g[4][4] = {{9, -15, -8, -17}, {-15, 10, -15, -4}, {-4, -15, 10, -15}, {-17, -8, -15, 9}};
static const int8_t TAB_25p43g[4][4] = {{8, -13, -6, -15}, {-13, 9, -13, -2}, {-2, -13, 9, -13}, {-15, -6, -13, 8}};

int
oracle_get_scoring(const char *name, int rs_match, int rs_mismatch, int rs_gap,
                   oracle_scoring_t *out)
{
  const int8_t (*tab)[4] = NULL;
  int a, b;
  memset(out, 0, sizeof(*out));
  if (strcmp(name, "repeatscout") == 0)
  {
    /* score_system.c:91-172: gapopen 0, gapextn = gap, every non-identical pair (including N
     * and the soft-masked a,c,g,t columns) scores `mismatch`. */
    for (a = 0; a < 4; a++)
      for (b = 0; b < ORACLE_NCODES; b++)
        out->matrix[a][b] = (a == b) ? rs_match : rs_mismatch;
    out->gapopen = 0;
    out->gapextn = rs_gap;
    return 0;
  }
  if (strcmp(name, "14p43g") == 0) { tab = TAB_14p43g; out->gapopen = -33; out->gapextn = -7; }
  else if (strcmp(name, "18p43g") == 0) { tab = TAB_18p43g; out->gapopen = -30; out->gapextn = -6; }
  else if (strcmp(name, "20p43g") == 0) { tab = TAB_20p43g; out->gapopen = -28; out->gapextn = -5; }
  else if (strcmp(name, "25p43g") == 0) { tab = TAB_25p43g; out->gapopen = -25; out->gapextn = -5; }
  else return -1;
  for (a = 0; a < 4; a++)
  {
    for (b = 0; b < 4; b++)
      out->matrix[a][b] = tab[a][b];
    /* soft-masked columns and N all score -1 (score_system.c:384-395) */
    for (b = 4; b < ORACLE_NCODES; b++)
      out->matrix[a][b] = -1;
  }
  return 0;
}

int
oracle_seq_code(int v)
{
  return (v >= 0 && v <= 7) ? v : 8;
}

int
oracle_compl_code(int c)
{
  /* sequence.c:1140-1160: A<->T, C<->G, a<->t, c<->g, everything else -> N */
  if (c >= 0 && c <= 3) return 3 - c;
  if (c >= 4 && c <= 7) return 11 - c;
  return 8;
}

/* ------------------------------------------------------------------ one band row --------- */

/* bnw_extend.c:778-788: first base outside the core in the direction of travel */
static int64_t
start_pos(int direction, const oracle_core_t *c)
{
  if (direction)
    return c->orient ? c->right_seq_pos - 1 : c->right_seq_pos + 1;
  return c->orient ? c->left_seq_pos + 1 : c->left_seq_pos - 1;
}

int
oracle_nw_row(int direction, int row_idx, int cons_base, const oracle_core_t *core,
              const int32_t *ps, const int32_t *pg, int32_t *cs, int32_t *cg,
              const uint8_t *sequence, const oracle_scoring_t *sc, int B, int *best_idx)
{
  const int go = sc->gapopen, ge = sc->gapextn;
  const int64_t start = start_pos(direction, core);
  /* bnw_extend.c:824-835: the sequence walks down when direction == orient */
  const int sgn = (direction == core->orient) ? -1 : 1;
  int best = ORACLE_NEG;
  int offset;

  for (offset = -B; offset <= B; offset++)
  {
    const int k = offset + B;
    const int64_t pos = start + (int64_t) sgn * (offset + row_idx);
    int sub, ins, del, gap, cell;
    /* the reference's three unsigned tests (bnw_extend.c:843-868) == this signed test */
    if (pos >= core->lower_seq_bound && pos <= core->upper_seq_bound)
    {
      int code = oracle_seq_code(sequence[pos]);
      int t, m;
      if (core->orient)
        code = oracle_compl_code(code);
      m = sc->matrix[cons_base][code];
      del = ORACLE_NEG;
      if (offset < B)
      {
        t = ps[k + 1] + go + ge; if (t > del) del = t;
        t = pg[k + 1] + ge;      if (t > del) del = t;
      }
      sub = ORACLE_NEG;
      t = ps[k] + m; if (t > sub) sub = t;
      t = pg[k] + m; if (t > sub) sub = t;
      ins = ORACLE_NEG;
      if (offset > -B)
      {
        t = cs[k - 1] + go + ge; if (t > ins) ins = t;
        t = cg[k - 1] + ge;      if (t > ins) ins = t;
      }
    }
    else if (offset < 0 && row_idx < B)
      sub = ins = del = go + (row_idx + 1) * ge;      /* bnw_extend.c:990-994 */
    else
      sub = ins = del = ORACLE_OOB;                   /* bnw_extend.c:995-1002 */

    cs[k] = sub;
    gap = (ins > del) ? ins : del;
    cg[k] = gap;
    cell = (gap > sub) ? gap : sub;
    if (cell > best)
    {
      best = cell;
      *best_idx = row_idx + offset;
    }
  }
  return best;
}

/* ------------------------------------------------------------------ column loop ---------- */

int
oracle_extend(int direction, int n, const oracle_core_t *cores, const uint8_t *sequence,
              const oracle_scoring_t *sc, int B, int cappenalty, int minimprovement, int L,
              int when_to_stop, uint8_t *master_out, int32_t *ext_len_out,
              int32_t *score_add_out, oracle_result_t *res, int32_t *trace_out)
{
  const int W = 2 * B + 1;
  const int go = sc->gapopen, ge = sc->gapextn;
  /* rolling state: committed previous row per instance */
  int32_t *ps = malloc((size_t) n * W * sizeof(int32_t));
  int32_t *pg = malloc((size_t) n * W * sizeof(int32_t));
  int32_t *high = calloc(n, sizeof(int32_t)), *pos = calloc(n, sizeof(int32_t));
  int32_t *thigh = calloc(n, sizeof(int32_t)), *tpos = calloc(n, sizeof(int32_t));
  int32_t *rowbest = malloc((size_t) n * 4 * sizeof(int32_t));
  int32_t *rowidx = malloc((size_t) n * 4 * sizeof(int32_t));
  /* candidate rows of every instance for the current column (commit = select) */
  int32_t *alls = malloc((size_t) n * 4 * W * sizeof(int32_t));
  int32_t *allg = malloc((size_t) n * 4 * W * sizeof(int32_t));
  int i, k, a, row_idx, n_ext = 0;
  int max_row = -1;
  uint32_t max_ext = 0;     /* sums wrap mod 2^32 exactly like the compiled reference's ints */
  int stop_row = -1;

  /* boundary row (ram_extend.c:909-946) */
  for (i = 0; i < n; i++)
    for (k = 0; k < W; k++)
    {
      int o = k - B;
      int v = (o == 0) ? 0 : (o < 0 ? -o : o) * ge + go;
      ps[(size_t) i * W + k] = v;
      pg[(size_t) i * W + k] = v;
    }
  for (i = 0; i < n; i++)
    if (direction ? cores[i].right_extendable : cores[i].left_extendable)
      n_ext++;

  for (row_idx = 0; row_idx < L; row_idx++)
  {
    uint32_t sums[4] = {0, 0, 0, 0};
    int32_t curr = 0;
    int besta = 0;
    for (i = 0; i < n; i++)
    {
      if (!(direction ? cores[i].right_extendable : cores[i].left_extendable))
        continue;
      for (a = 0; a < 4; a++)
      {
        int idx = 0;
        int32_t *as = alls + ((size_t) i * 4 + a) * W, *ag = allg + ((size_t) i * 4 + a) * W;
        int best = oracle_nw_row(direction, row_idx, a, &cores[i], ps + (size_t) i * W,
                                 pg + (size_t) i * W, as, ag, sequence, sc, B, &idx);
        int c = best < 0 ? 0 : best;                     /* ram_extend.c:1042-1043 */
        int cap = high[i] + cappenalty;
        rowbest[i * 4 + a] = best;
        rowidx[i * 4 + a] = idx;
        sums[a] += (uint32_t) ((c >= cap) ? c : cap);    /* ram_extend.c:1052-1062 */
      }
    }
    for (a = 0; a < 4; a++)                               /* ram_extend.c:1081-1085 */
      if ((int32_t) sums[a] > curr)
      {
        curr = (int32_t) sums[a];
        besta = a;
      }
    master_out[row_idx] = (uint8_t) besta;

    /* commit the chosen candidate's row (== the reference's 5th recompute pass, 1105-1150) */
    for (i = 0; i < n; i++)
    {
      if (!(direction ? cores[i].right_extendable : cores[i].left_extendable))
        continue;
      memcpy(ps + (size_t) i * W, alls + ((size_t) i * 4 + besta) * W, W * sizeof(int32_t));
      memcpy(pg + (size_t) i * W, allg + ((size_t) i * 4 + besta) * W, W * sizeof(int32_t));
      if (rowbest[i * 4 + besta] > high[i])
      {
        high[i] = rowbest[i * 4 + besta];
        pos[i] = rowidx[i * 4 + besta];
      }
    }
    /* ram_extend.c:1194-1208 */
    {
      int d = max_row - row_idx;
      if (d < 0) d = -d;
      if (curr >= (int32_t) (max_ext + (uint32_t) (d * minimprovement)))
      {
        max_row = row_idx;
        max_ext = (uint32_t) curr;
        memcpy(thigh, high, n * sizeof(int32_t));
        memcpy(tpos, pos, n * sizeof(int32_t));
      }
    }
    if (trace_out)
    {
      int32_t *t = trace_out + (size_t) row_idx * 6;
      t[0] = (int32_t) sums[0]; t[1] = (int32_t) sums[1];
      t[2] = (int32_t) sums[2]; t[3] = (int32_t) sums[3];
      t[4] = besta; t[5] = max_row;
    }
    if (row_idx - max_row >= when_to_stop)                /* ram_extend.c:1216-1223 */
    {
      stop_row = row_idx;
      break;
    }
  }

  for (i = 0; i < n; i++)                                 /* ram_extend.c:1234-1247 */
  {
    ext_len_out[i] = 0;
    score_add_out[i] = 0;
    if (thigh[i] > 0 && tpos[i] >= 0)
    {
      ext_len_out[i] = tpos[i] + 1;
      score_add_out[i] = thigh[i];
    }
  }
  res->extended_bp = max_row + 1;
  res->rows_processed = (stop_row >= 0) ? stop_row + 1 : (L > 0 ? L : 0);
  res->stop_row = stop_row;
  res->hit_limit = (stop_row >= 0 && stop_row == L - 1) ? 1 : 0;   /* ram_extend.c:1225 */
  res->max_score = (int32_t) max_ext;
  res->n_extendable = n_ext;

  free(ps); free(pg); free(high); free(pos); free(thigh); free(tpos);
  free(rowbest); free(rowidx); free(alls); free(allg);
  return 0;
}

double
oracle_cell_updates(const oracle_result_t *res, int bandwidth)
{
  return (double) res->rows_processed * res->n_extendable * (2.0 * bandwidth + 1.0) * 4.0;
}
