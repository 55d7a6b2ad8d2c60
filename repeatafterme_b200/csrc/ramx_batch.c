/*
 * ramx_batch.c -- RAMExtendBatch: many RAMExtend runs in one process, on one CUDA context per GPU.
 *
 *     RAMExtendBatch [-gpus N] [-procs P] [-chunk K] <manifest>
 *
 * Replaces what util/extend-stk.pl:365-371 (and a RepeatModeler round) does with one
 * `system("RAMExtend ...")` per family. Every manifest line is the argument list of one RAMExtend
 * run (everything after the program name), optionally preceded by `stdout=<file>`:
 *
 *     stdout=fam1.log -twobit ce10.2bit -ranges fam1.tsv -L 20000 -bandwidth 40 -matrix 14p43g \
 *                     -minimprovement 30 -outtsv fam1.out.tsv -outfa fam1.out.fa -cons fam1.cons.fa -vvv
 *
 * and produces exactly the files and the stdout a separate run of RAMExtend with those arguments
 * produces: flag parsing, the loader's rules, overlap avoidance, consensus / report / -outtsv / -outfa
 * are the reference's own main() (ram_extend.c:217-790, compiled out of tree, renamed ramref_main).
 * What differs is when the DP runs. Families are processed in chunks, three phases per chunk:
 *   1. record : main() of every family runs up to its first extend_alignment() call with stdout
 *               discarded; the call's flattened inputs are captured, the library goes to the device
 *               (2 bits per base), both passes are created as a linked pair, and main() is left by
 *               longjmp.
 *   2. run    : ONE ramcu_batch_run over all 2 x K column loops of the chunk -- the families share
 *               the GPU instead of taking turns -- and one fetch per job.
 *   3. replay : main() of every family runs again, to the end, with stdout going to its file;
 *               extend_alignment() is answered from the results. The left pass was computed with the
 *               bounds as they were before overlap avoidance (ram_extend.c:445-499): if that step
 *               narrowed a bound, the left call's inputs differ from the recorded ones and the pass is
 *               run again on its own, exactly as the single-family CLI does.
 * -outmat runs and anything else the hook declines go through the boundary's own single-family path.
 * With -gpus N the manifest is split over N worker processes (longest first by ranges-file size x band
 * width), one GPU each; families are independent, nothing is exchanged. -procs P starts P workers per GPU:
 * the record and replay phases are the reference's single-threaded host code (loader, report, output files),
 * so for many small families the host, not the GPU, sets the pace; P processes overlap their host phases and
 * take turns on the device (separate contexts are time-sliced, a kernel always has the device to itself).
 * An error that makes the reference's main() call exit() ends the worker that hit it.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "common.h"     /* reference headers (-I$(REF)): struct sequenceLibrary */
#include "sequence.h"

#include "ramcu.h"
#include "ramx_boundary.h"
#include "ramx_loader.h"

int ramref_main(int argc, char **argv);   /* the reference's main(), renamed at link time */

/* globals of ram_extend.c that flag parsing only sets when the flag is present (ram_extend.c:235-245) */
extern char *SEQUENCE_FILE, *MATRIX_FILE, *OUTTSV_FILE, *OUTFA_FILE, *CONS_FILE, *RANGES_FILE, *OUTMAT_FILE;
extern int flanking;

struct family
{
  int argc;
  char **argv;          /* argv[0] = "RAMExtend" */
  char *stdout_path;    /* NULL: the batch's own stdout */
  double cost;
  /* recorded */
  int recorded, n, L;
  ramcu_core_t *cores;
  ramcu_scoring_t sc;
  ramcu_params_t pr;
  ramcu_library *lib;
  ramcu_job *job[2];    /* right, left */
  /* results */
  uint8_t *cons[2];
  int32_t *ext_len[2], *score_add[2];
  ramcu_result_t res[2];
};

static struct family *g_fam = NULL;   /* the family whose main() is running */
static int g_mode = 0;                /* 1 record, 2 replay */
static jmp_buf g_leave;

static double
now_s(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

static void
fail(const char *what)
{
  fflush(stdout);
  fprintf(stderr, "RAMExtendBatch: %s: %s\n", what, ramcu_last_error());
  exit(1);
}

static int
hook(int direction, int n, const ramcu_core_t *cores, const ramcu_scoring_t *sc, const ramcu_params_t *pr,
     struct sequenceLibrary *seqLib, uint8_t *cons, int32_t *ext_len, int32_t *score_add, ramcu_result_t *res)
{
  struct family *f = g_fam;
  if (f == NULL)
    return 0;
  if (g_mode == 1)
  {
    if (direction != 1)
      return 0;   /* main() always extends right first */
    ramcu_ctx *ctx = ramx_boundary_context();
    const struct ramx_twobit_source *src = ramx_loader_source(seqLib->sequence, seqLib->length);
    f->n = n;
    f->L = pr->L;
    f->sc = *sc;
    f->pr = *pr;
    f->cores = (ramcu_core_t *) malloc((size_t) (n > 0 ? n : 1) * sizeof(ramcu_core_t));
    memcpy(f->cores, cores, (size_t) n * sizeof(ramcu_core_t));
    if (src == NULL)
      return 0;   /* not our loader's library: let the boundary handle this family on its own */
    if (ramcu_library_from_twobit(ctx, src->blob, src->blob_bytes, src->n_windows, (const ramcu_window_t *) src->windows,
                                  src->n_nranges, src->n_start, src->n_len, src->total, &f->lib) != RAMCU_OK)
      fail("ramcu_library_from_twobit");
    if (ramcu_job_create_pair(ctx, f->lib, n, f->cores, &f->sc, &f->pr, &f->job[0], &f->job[1]) != RAMCU_OK)
      fail("ramcu_job_create_pair");
    f->recorded = 1;
    /* this run of main() ends here; the buffers of the call are the boundary's plain mallocs */
    free((void *) cores);
    free(cons);
    free(ext_len);
    free(score_add);
    longjmp(g_leave, 1);
  }
  if (g_mode == 2 && f->recorded)
  {
    const int d = direction ? 0 : 1;
    if (n != f->n || memcmp(pr, &f->pr, sizeof(*pr)) != 0 || memcmp(sc, &f->sc, sizeof(*sc)) != 0 ||
        memcmp(cores, f->cores, (size_t) n * sizeof(ramcu_core_t)) != 0)
      return 0;   /* e.g. overlap avoidance narrowed a bound before the left pass */
    memcpy(cons, f->cons[d], (size_t) (f->L > 0 ? f->L : 1));
    memcpy(ext_len, f->ext_len[d], (size_t) n * sizeof(int32_t));
    memcpy(score_add, f->score_add[d], (size_t) n * sizeof(int32_t));
    *res = f->res[d];
    return 1;
  }
  return 0;
}

static void
run_main(struct family *f, const char *out_path, int append)
{
  fflush(stdout);
  const int saved = dup(1);
  if (out_path != NULL)
  {
    const int fd = open(out_path, O_WRONLY | O_CREAT | (append ? O_APPEND : O_TRUNC), 0644);
    if (fd < 0)
    {
      fprintf(stderr, "RAMExtendBatch: cannot write %s\n", out_path);
      exit(1);
    }
    dup2(fd, 1);
    close(fd);
  }
  SEQUENCE_FILE = MATRIX_FILE = OUTTSV_FILE = OUTFA_FILE = CONS_FILE = RANGES_FILE = OUTMAT_FILE = NULL;
  flanking = 0;
  ramx_boundary_forget_library();
  g_fam = f;
  if (setjmp(g_leave) == 0)
    ramref_main(f->argc, f->argv);
  g_fam = NULL;
  fflush(stdout);
  dup2(saved, 1);
  close(saved);
  ramx_loader_release_last();
}

/* ---------------------------------------------------------------------------------------------
 * manifest
 * ------------------------------------------------------------------------------------------- */
static struct family *
read_manifest(const char *path, int *n_out)
{
  FILE *fh = fopen(path, "r");
  if (!fh)
  {
    fprintf(stderr, "RAMExtendBatch: cannot read %s\n", path);
    exit(1);
  }
  int cap = 64, n = 0;
  struct family *fams = (struct family *) calloc((size_t) cap, sizeof(*fams));
  char *line = NULL;
  size_t len = 0;
  while (getline(&line, &len, fh) > 0)
  {
    char *p = line;
    while (*p == ' ' || *p == '\t')
      ++p;
    if (*p == '#' || *p == '\n' || *p == 0)
      continue;
    if (n == cap)
    {
      cap *= 2;
      fams = (struct family *) realloc(fams, (size_t) cap * sizeof(*fams));
      memset(fams + n, 0, (size_t) (cap - n) * sizeof(*fams));
    }
    struct family *f = &fams[n];
    char *text = strdup(p);
    int ac = 1, acap = 16;
    f->argv = (char **) malloc((size_t) acap * sizeof(char *));
    f->argv[0] = (char *) "RAMExtend";
    for (char *tok = strtok(text, " \t\r\n"); tok != NULL; tok = strtok(NULL, " \t\r\n"))
    {
      if (ac == 1 && strncmp(tok, "stdout=", 7) == 0 && f->stdout_path == NULL)
      {
        f->stdout_path = tok + 7;
        continue;
      }
      if (ac + 2 > acap)
      {
        acap *= 2;
        f->argv = (char **) realloc(f->argv, (size_t) acap * sizeof(char *));
      }
      f->argv[ac++] = tok;
    }
    f->argv[ac] = NULL;
    f->argc = ac;
    /* cost estimate for the split over GPUs: size of the ranges file x band width */
    double ranges_bytes = 1.0, band = 29.0;
    for (int i = 1; i + 1 < ac; ++i)
    {
      struct stat st;
      if (strcmp(f->argv[i], "-ranges") == 0 && stat(f->argv[i + 1], &st) == 0)
        ranges_bytes = (double) st.st_size + 1.0;
      if (strcmp(f->argv[i], "-bandwidth") == 0)
        band = 2.0 * atof(f->argv[i + 1]) + 1.0;
    }
    f->cost = ranges_bytes * band;
    ++n;
  }
  free(line);
  fclose(fh);
  *n_out = n;
  return fams;
}

/* ---------------------------------------------------------------------------------------------
 * one worker: its families, in chunks
 * ------------------------------------------------------------------------------------------- */
static void
process(struct family **fams, int n, int chunk)
{
  ramcu_ctx *ctx = ramx_boundary_context();
  ramx_extend_hook = hook;
  const double t_begin = now_s();
  double t_rec = 0, t_run = 0, t_rep = 0;
  for (int c0 = 0; c0 < n; c0 += chunk)
  {
    const int c1 = c0 + chunk < n ? c0 + chunk : n;
    double t0 = now_s();
    g_mode = 1;
    for (int i = c0; i < c1; ++i)
      run_main(fams[i], "/dev/null", 0);
    t_rec += now_s() - t0;
    t0 = now_s();
    /* all column loops of the chunk share the device */
    int n_jobs = 0;
    ramcu_job **jobs = (ramcu_job **) malloc((size_t) (2 * (c1 - c0) + 1) * sizeof(ramcu_job *));
    for (int i = c0; i < c1; ++i)
      if (fams[i]->recorded)
      {
        jobs[n_jobs++] = fams[i]->job[0];
        jobs[n_jobs++] = fams[i]->job[1];
      }
    if (n_jobs > 0 && ramcu_batch_run(ctx, n_jobs, jobs) != RAMCU_OK)
      fail("ramcu_batch_run");
    free(jobs);
    for (int i = c0; i < c1; ++i)
    {
      struct family *f = fams[i];
      if (!f->recorded)
        continue;
      for (int d = 0; d < 2; ++d)
      {
        f->cons[d] = (uint8_t *) malloc((size_t) (f->L > 0 ? f->L : 1));
        f->ext_len[d] = (int32_t *) calloc((size_t) (f->n > 0 ? f->n : 1), sizeof(int32_t));
        f->score_add[d] = (int32_t *) calloc((size_t) (f->n > 0 ? f->n : 1), sizeof(int32_t));
        if (ramcu_job_fetch(ctx, f->job[d], f->cons[d], f->ext_len[d], f->score_add[d], &f->res[d]) != RAMCU_OK)
          fail("ramcu_job_fetch");
        ramcu_job_free(ctx, f->job[d]);
      }
      ramcu_library_free(ctx, f->lib);
    }
    t_run += now_s() - t0;
    t0 = now_s();
    g_mode = 2;
    for (int i = c0; i < c1; ++i)
    {
      struct family *f = fams[i];
      run_main(f, f->stdout_path, 0);
      for (int d = 0; d < 2; ++d)
      {
        free(f->cons[d]);
        free(f->ext_len[d]);
        free(f->score_add[d]);
      }
      free(f->cores);
    }
    t_rep += now_s() - t0;
  }
  const double dt = now_s() - t_begin;
  fprintf(stderr, "RAMExtendBatch: %d families in %.3f s (%.1f families/s): record %.3f s, GPU batches + fetch %.3f s, "
          "replay (reports and output files) %.3f s\n", n, dt, n / (dt > 0 ? dt : 1), t_rec, t_run, t_rep);
}

int
main(int argc, char **argv)
{
  int gpus = 1, procs = 1, chunk = 512;
  const char *manifest = NULL;
  for (int i = 1; i < argc; ++i)
  {
    if (strcmp(argv[i], "-gpus") == 0 && i + 1 < argc)
      gpus = atoi(argv[++i]);
    else if (strcmp(argv[i], "-procs") == 0 && i + 1 < argc)
      procs = atoi(argv[++i]);
    else if (strcmp(argv[i], "-chunk") == 0 && i + 1 < argc)
      chunk = atoi(argv[++i]);
    else
      manifest = argv[i];
  }
  if (manifest == NULL || gpus < 1 || procs < 1 || chunk < 1)
  {
    fprintf(stderr, "usage: RAMExtendBatch [-gpus N] [-procs P] [-chunk K] <manifest>\n"
            "  manifest: one RAMExtend argument list per line, optionally preceded by stdout=<file>\n");
    return 1;
  }
  const int workers = gpus * procs;   /* worker w runs on GPU w % gpus */
  int n = 0;
  struct family *fams = read_manifest(manifest, &n);
  /* longest-processing-time-first split over the workers */
  int *order = (int *) malloc((size_t) (n > 0 ? n : 1) * sizeof(int));
  for (int i = 0; i < n; ++i)
    order[i] = i;
  for (int i = 1; i < n; ++i)   /* insertion sort by decreasing cost, stable */
  {
    const int v = order[i];
    int j = i - 1;
    while (j >= 0 && fams[order[j]].cost < fams[v].cost)
    {
      order[j + 1] = order[j];
      --j;
    }
    order[j + 1] = v;
  }
  int *owner = (int *) malloc((size_t) (n > 0 ? n : 1) * sizeof(int));
  double *load = (double *) calloc((size_t) workers, sizeof(double));
  for (int k = 0; k < n; ++k)
  {
    int best = 0;
    for (int g = 1; g < workers; ++g)
      if (load[g] < load[best])
        best = g;
    owner[order[k]] = best;
    load[best] += fams[order[k]].cost;
  }
  int status = 0;
  for (int g = 0; g < workers; ++g)
  {
    pid_t pid = workers > 1 ? fork() : 0;
    if (pid == 0)
    {
      if (gpus > 1)
      {
        char dev[16];
        snprintf(dev, sizeof(dev), "%d", g % gpus);
        setenv("RAMCU_DEVICE", dev, 1);
      }
      int m = 0;
      struct family **mine = (struct family **) malloc((size_t) (n > 0 ? n : 1) * sizeof(*mine));
      for (int i = 0; i < n; ++i)   /* manifest order inside a worker */
        if (owner[i] == g)
          mine[m++] = &fams[i];
      if (m > 0)
        process(mine, m, chunk);
      if (workers > 1)
        _exit(0);
    }
    else if (pid < 0)
    {
      fprintf(stderr, "RAMExtendBatch: fork failed\n");
      return 1;
    }
  }
  if (workers > 1)
  {
    int st;
    while (wait(&st) > 0)
      if (!WIFEXITED(st) || WEXITSTATUS(st) != 0)
        status = 1;
  }
  return status;
}
