/*
 * ramx_loader.c -- host-side C: a strong definition of the reference's flank loader
 *
 *     struct sequenceLibrary *loadSequenceSubsetMinimal(char *twoBitName, char *rangeBEDName,
 *                                                       struct coreAlignment **core_align,
 *                                                       int *num_cores, int max_flanking_bp);
 *
 * (sequence.h, defined in sequence.c:505-923) for the product CLI: the symbol is weakened in the
 * reference's sequence.o and main() (ram_extend.c:369) calls this one. Same outputs -- the
 * byte-per-base seqLib (sequence / identifiers / boundaries / offsets / length / count), the core
 * list in the loader's sorted order with the same bounds and flags, the same WARNING lines -- built
 * differently:
 *   - the .2bit file is mapped once and its index parsed once into a sorted table (the reference
 *     seeks and re-reads the record header for every core: twoBitSeqSize + twoBitReadSeqFrag,
 *     kentsrc/twoBitNew.c:487-664);
 *   - the window rules of sequence.c:625-743 are applied to the sorted ranges in one pass, the
 *     total size is known before anything is decoded, and the library is ONE allocation (the
 *     reference reallocs four arrays per core, sequence.c:763-797);
 *   - windows are decoded straight from the packed DNA to the byte codes (A0 C1 G2 T3, N 99) with a
 *     256-entry table, four bases per look-up, instead of 2bit -> ASCII -> toUpperN -> codes;
 *   - the windows themselves (packed bytes as the file stores them + N runs) are handed to
 *     ramx_boundary.c, so that the device decodes its library from 2 bits per base
 *     (ramcu_library_from_twobit) and never sees the byte expansion; the bytes above exist only
 *     for the reference's own report / -outfa / overlap code.
 *
 * Facts of the format: kentsrc/twoBitNew.c:250-290 (header, index), :571-582 (packed DNA, T0 C1 A2
 * G3, first base in the top bits), :597-613 (N blocks). Soft-mask blocks are skipped: RAMExtend
 * upper-cases everything (sequence.c:759).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "common.h"    /* struct coreAlignment, enum CoreBoundFlag (reference, -I$(REF)) */
#include "sequence.h"  /* struct sequenceLibrary */
#include "score_system.h"

#include "ramx_loader.h"

static void
die(const char *fmt, const char *a, const char *b)
{
  fflush(stdout);
  fprintf(stderr, fmt, a, b);
  fprintf(stderr, "\n");
  exit(255);   /* errAbort() of the kent library exits with -1 */
}

/* ---------------------------------------------------------------------------------------------
 * BED-6 ranges (readBEDRanges, sequence.c:942-976; kentsrc/linefile.c:690-706, common.c:2101)
 * ------------------------------------------------------------------------------------------- */
struct range
{
  char *name;
  long start, end;
  int left_flag, right_flag;   /* atoi(BED name) == 1, atoi(BED score) == 1 */
  int left_zero, right_zero;   /* atoi(...) == 0: the neighbour tests of sequence.c:635,662,693,721 */
  int reverse;                 /* strcmp(strand, "-") == 0 */
  int order;                   /* position in the file: the sort below is stable, as glibc's qsort is */
  /* filled by the window pass */
  long win_start, win_end;
  int lower_flank_len;
  int lower_flag, upper_flag;
  int record;                  /* index into the .2bit table */
};

/* everything the last call allocated (main() never frees; the batch front-end does, through
 * ramx_loader_release_last(), once a run of main() is over) */
static struct
{
  struct sequenceLibrary *lib;
  struct coreAlignment *cores;
  struct range *ranges;
  char *ranges_text;
  int n;
} g_last;

static struct range *
read_ranges(const char *path, int *n_out)
{
  FILE *fh = fopen(path, "rb");
  if (!fh)
    die("Couldn't open %s , %s", path, strerror(errno));
  long size = -1;
  if (fseek(fh, 0, SEEK_END) == 0)
    size = ftell(fh);
  if (size < 0 || fseek(fh, 0, SEEK_SET) != 0)
    die("Couldn't read %s: %s", path, "not a regular file");
  char *buf = (char *) malloc((size_t) size + 2);
  if (!buf || fread(buf, 1, (size_t) size, fh) != (size_t) size)
    die("Couldn't read %s%s", path, "");
  fclose(fh);
  buf[size] = '\n';
  buf[size + 1] = 0;
  int cap = 1024, n = 0;
  struct range *r = (struct range *) malloc(cap * sizeof(*r));
  if (!r)
    die("Out of memory reading %s%s", path, "");
  char *line = buf, *end = buf + size;
  while (line < end)
  {
    char *nl = memchr(line, '\n', (size_t) (end - line) + 1);
    *nl = 0;
    if (nl > line && nl[-1] == '\r')
      nl[-1] = 0;
    if (line[0] != '#' && line[0] != 0)
    {
      /* chopByChar: every tab ends a field, at most six fields */
      char *f[6];
      int nf = 0;
      char *p = line;
      while (nf < 6)
      {
        f[nf++] = p;
        char *t = strchr(p, '\t');
        if (!t)
          break;
        *t = 0;
        p = t + 1;
      }
      if (nf < 6 || !(f[5][0] == '+' || f[5][0] == '-'))
      {
        printf("Error: ranges file does not appear to be in the correct format!\n");
        exit(1);
      }
      if (n == cap)
      {
        cap *= 2;
        r = (struct range *) realloc(r, cap * sizeof(*r));
        if (!r)
          die("Out of memory reading %s%s", path, "");
      }
      char *e;
      memset(&r[n], 0, sizeof(r[n]));
      r[n].name = f[0];
      r[n].start = strtol(f[1], &e, 0);
      r[n].end = strtol(f[2], &e, 0);
      r[n].left_flag = atoi(f[3]) == 1;
      r[n].right_flag = atoi(f[4]) == 1;
      r[n].left_zero = atoi(f[3]) == 0;
      r[n].right_zero = atoi(f[4]) == 0;
      r[n].reverse = strcmp(f[5], "-") == 0;
      r[n].order = n;
      n++;
    }
    line = nl + 1;
  }
  *n_out = n;
  g_last.ranges_text = buf;
  return r;   /* names point into buf, which lives as long as the ranges */
}

/* bedNameStartRevEndCmp (sequence.c:493-503): name, start ascending, end descending; the int
 * differences of the reference are kept */
static int
range_cmp(const void *va, const void *vb)
{
  const struct range *a = (const struct range *) va, *b = (const struct range *) vb;
  int diff = strcmp(a->name, b->name);
  if (diff == 0)
    diff = (int) a->start - (int) b->start;
  if (diff == 0)
    diff = (int) b->end - (int) a->end;
  if (diff == 0)
    diff = a->order - b->order;
  return diff;
}

/* ---------------------------------------------------------------------------------------------
 * .2bit file: mapped, index parsed once
 * ------------------------------------------------------------------------------------------- */
struct record
{
  const char *name;
  int name_len;
  uint64_t offset;        /* of the record */
  uint32_t size;          /* bases */
  uint32_t n_blocks;
  uint64_t nstart_off;    /* file offsets of the N block arrays */
  uint64_t nsize_off;
  uint64_t dna_off;       /* file offset of the packed DNA */
  int parsed;
};

struct twobit
{
  const uint8_t *map;
  uint64_t bytes;
  int swapped, version;
  int n_records;
  struct record *rec;     /* sorted by name */
};

static uint32_t
rd32(const struct twobit *t, uint64_t off)
{
  if (off + 4 > t->bytes)
    die("%s: truncated .2bit file%s", "loadSequenceSubsetMinimal", "");
  uint32_t v;
  memcpy(&v, t->map + off, 4);
  return t->swapped ? __builtin_bswap32(v) : v;
}
static uint64_t
rd64(const struct twobit *t, uint64_t off)
{
  if (off + 8 > t->bytes)
    die("%s: truncated .2bit file%s", "loadSequenceSubsetMinimal", "");
  uint64_t v;
  memcpy(&v, t->map + off, 8);
  return t->swapped ? __builtin_bswap64(v) : v;
}

static int
record_cmp(const void *va, const void *vb)
{
  const struct record *a = (const struct record *) va, *b = (const struct record *) vb;
  const int n = a->name_len < b->name_len ? a->name_len : b->name_len;
  int d = memcmp(a->name, b->name, (size_t) n);
  if (d == 0)
    d = a->name_len - b->name_len;
  return d;
}

static void
twobit_open(struct twobit *t, const char *path)
{
  int fd = open(path, O_RDONLY);
  if (fd < 0)
    die("Can't open %s to read: %s", path, strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 16)
    die("%s is not a twoBit file%s", path, "");
  t->bytes = (uint64_t) st.st_size;
  t->map = (const uint8_t *) mmap(NULL, t->bytes, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (t->map == MAP_FAILED)
    die("Can't map %s: %s", path, strerror(errno));
  uint32_t sig;
  memcpy(&sig, t->map, 4);
  if (sig == 0x1A412743u)
    t->swapped = 0;
  else if (sig == 0x4327411Au)
    t->swapped = 1;
  else
    die("%s doesn't have a valid twoBitSig%s", path, "");
  t->version = (int) rd32(t, 4);
  if (t->version != 0 && t->version != 1)
    die("Can only handle version 0 or version 1 of this file. This is version %s%s", "?", "");
  t->n_records = (int) rd32(t, 8);
  if (t->n_records < 0)
    die("%s: corrupt sequence count%s", path, "");
  t->rec = (struct record *) calloc((size_t) (t->n_records > 0 ? t->n_records : 1), sizeof(struct record));
  if (!t->rec)
    die("%s: out of memory for the index%s", path, "");
  uint64_t p = 16;
  for (int i = 0; i < t->n_records; ++i)
  {
    if (p + 1 > t->bytes)
      die("%s: truncated index%s", path, "");
    const int len = t->map[p];
    t->rec[i].name = (const char *) t->map + p + 1;
    t->rec[i].name_len = len;
    p += 1 + (uint64_t) len;
    if (t->version == 1)
    {
      t->rec[i].offset = rd64(t, p);
      p += 8;
    }
    else
    {
      t->rec[i].offset = rd32(t, p);
      p += 4;
    }
  }
  qsort(t->rec, (size_t) t->n_records, sizeof(struct record), record_cmp);
}

static int
twobit_find(struct twobit *t, const char *name, const char *path)
{
  struct record key;
  key.name = name;
  key.name_len = (int) strlen(name);
  struct record *r = (struct record *) bsearch(&key, t->rec, (size_t) t->n_records, sizeof(struct record), record_cmp);
  if (!r)
    die("%s is not in %s", name, path);
  if (!r->parsed)
  {
    /* dnaSize, nBlockCount, nStarts[], nSizes[], maskBlockCount, maskStarts[], maskSizes[], reserved */
    uint64_t p = r->offset;
    r->size = rd32(t, p);
    r->n_blocks = rd32(t, p + 4);
    r->nstart_off = p + 8;
    r->nsize_off = r->nstart_off + 4ull * r->n_blocks;
    p = r->nsize_off + 4ull * r->n_blocks;
    const uint32_t mask_blocks = rd32(t, p);
    p += 4 + 8ull * mask_blocks + 4;
    r->dna_off = p;
    if (r->dna_off + ((uint64_t) r->size + 3) / 4 > t->bytes)
      die("%s: truncated record %s", path, name);
    r->parsed = 1;
  }
  return (int) (r - t->rec);
}

/* packed byte -> four byte codes (A0 C1 G2 T3); .2bit stores T0 C1 A2 G3, first base in the top bits */
static uint32_t g_quad[256];
static void
init_quad(void)
{
  static const uint8_t map[4] = {3, 1, 0, 2};
  for (int b = 0; b < 256; ++b)
    g_quad[b] = (uint32_t) map[(b >> 6) & 3] | ((uint32_t) map[(b >> 4) & 3] << 8) |
                ((uint32_t) map[(b >> 2) & 3] << 16) | ((uint32_t) map[b & 3] << 24);
}

static void
decode_window(const struct twobit *t, const struct record *rec, long start, long end, char *out)
{
  const uint8_t *dna = t->map + rec->dna_off;
  long p = start;
  while (p < end && (p & 3))   /* head */
  {
    *out++ = (char) ((g_quad[dna[p >> 2]] >> (8 * (p & 3))) & 0xff);
    ++p;
  }
  while (p + 4 <= end)
  {
    const uint32_t q = g_quad[dna[p >> 2]];
    memcpy(out, &q, 4);
    out += 4;
    p += 4;
  }
  while (p < end)              /* tail */
  {
    *out++ = (char) ((g_quad[dna[p >> 2]] >> (8 * (p & 3))) & 0xff);
    ++p;
  }
}

/* ---------------------------------------------------------------------------------------------
 * what the device needs: the windows as packed bytes + N runs (picked up by ramx_boundary.c)
 * ------------------------------------------------------------------------------------------- */
static struct ramx_twobit_source g_source;


const struct ramx_twobit_source *
ramx_loader_source(const char *sequence, uint64_t length)
{
  if (g_source.blob != NULL && g_source.sequence == sequence && g_source.total == length)
    return &g_source;
  return NULL;
}

const struct ramx_twobit_source *
ramx_loader_last_source(void)
{
  return g_source.blob != NULL ? &g_source : NULL;
}

static void
nrange_push(uint64_t start, uint64_t len)
{
  if (g_source.n_nranges == g_source.cap_nranges)
  {
    g_source.cap_nranges = g_source.cap_nranges ? 2 * g_source.cap_nranges : 256;
    g_source.n_start = (uint64_t *) realloc(g_source.n_start, (size_t) g_source.cap_nranges * sizeof(uint64_t));
    g_source.n_len = (uint64_t *) realloc(g_source.n_len, (size_t) g_source.cap_nranges * sizeof(uint64_t));
    if (!g_source.n_start || !g_source.n_len)
      die("%s: out of memory for the N runs%s", "loadSequenceSubsetMinimal", "");
  }
  g_source.n_start[g_source.n_nranges] = start;
  g_source.n_len[g_source.n_nranges] = len;
  g_source.n_nranges++;
}

/* ---------------------------------------------------------------------------------------------
 * the loader
 * ------------------------------------------------------------------------------------------- */
struct sequenceLibrary *
loadSequenceSubsetMinimal(char *twoBitName, char *rangeBEDName, struct coreAlignment **core_align,
                          int *num_cores, int max_flanking_bp)
{
  struct timespec t_begin;
  clock_gettime(CLOCK_MONOTONIC, &t_begin);
  int n = 0;
  struct range *rg = read_ranges(rangeBEDName, &n);
  struct twobit tb;
  memset(&tb, 0, sizeof(tb));
  twobit_open(&tb, twoBitName);
  init_quad();

  /* "Can't depend on the BED file being sorted" (sequence.c:533-535) */
  qsort(rg, (size_t) n, sizeof(*rg), range_cmp);

  /* ---- windows: one pass over the sorted ranges (sequence.c:538-743) ---------------------------- */
  uint64_t total = 0;
  for (int i = 0; i < n; ++i)
  {
    struct range *s = &rg[i];
    s->record = twobit_find(&tb, s->name, twoBitName);
    const int seq_size = (int) tb.rec[s->record].size;
    int lower_flank_len = 0;
    int flanking_start = (int) s->start, flanking_end = (int) s->end;
    int lower_flag = L_BOUNDARY, upper_flag = L_BOUNDARY;
    const struct range *prev = NULL, *next = NULL;
    int prev_dist = 0, next_dist = 0;
    if (i > 0 && strcmp(s->name, rg[i - 1].name) == 0)
    {
      const struct range *q = &rg[i - 1];
      if (s->start < q->end)
      {
        printf("WARNING: core sequences overlap  %s:%d-%d and previous %s:%d-%d\n", s->name, (int) s->start,
               (int) s->end, q->name, (int) q->start, (int) q->end);
        prev_dist = 0;
      }
      else
        prev_dist = (int) (s->start - q->end);
      prev = q;
    }
    if (i + 1 < n && strcmp(s->name, rg[i + 1].name) == 0)
    {
      const struct range *q = &rg[i + 1];
      if (q->start < s->end)
      {
        printf("WARNING: core sequences overlap  %s:%d-%d and next %s:%d-%d\n", s->name, (int) s->start,
               (int) s->end, q->name, (int) q->start, (int) q->end);
        next_dist = 0;
      }
      else
        next_dist = (int) (q->start - s->end);
      next = q;
    }
    /* The reference compares the strand STRINGS' addresses (s->strand == next->strand), which never
     * match, so "same strand" never applies: the whole gap goes to this core only when the neighbour
     * is not extendable towards it, else the gap is halved -- ceil()/floor() of an int division, i.e.
     * both truncate. The side a flag widens depends on the strand (sequence.c:616-743). */
    const int want_upper = s->reverse ? s->left_flag : s->right_flag;
    const int want_lower = s->reverse ? s->right_flag : s->left_flag;
    if (want_upper)
    {
      const int next_blocks = next ? (s->reverse ? next->left_zero : next->right_zero) : 0;
      if (next && next_dist <= max_flanking_bp)
      {
        flanking_end = (int) s->end + (next_blocks ? next_dist : next_dist / 2);
        upper_flag = CORE_BOUNDARY;
      }
      else if (s->end + max_flanking_bp < seq_size)
      {
        flanking_end = (int) s->end + max_flanking_bp;
        upper_flag = L_BOUNDARY;
      }
      else
      {
        flanking_end = seq_size;
        upper_flag = SEQ_BOUNDARY;
      }
    }
    if (want_lower)
    {
      const int prev_blocks = prev ? (s->reverse ? prev->right_zero : prev->left_zero) : 0;
      if (prev && prev_dist <= max_flanking_bp)
      {
        lower_flank_len = prev_blocks ? prev_dist : prev_dist / 2;
        flanking_start = (int) s->start - lower_flank_len;
        lower_flag = CORE_BOUNDARY;
      }
      else if (max_flanking_bp < s->start)
      {
        flanking_start = (int) s->start - max_flanking_bp;
        lower_flank_len = max_flanking_bp;
        lower_flag = L_BOUNDARY;
      }
      else
      {
        flanking_start = 0;
        lower_flank_len = (int) s->start;
        lower_flag = SEQ_BOUNDARY;
      }
    }
    /* twoBitReadSeqFragExt's range checks (kentsrc/twoBitNew.c:507-513) */
    int frag_end = flanking_end;
    if (frag_end == 0)
      frag_end = seq_size;
    if (frag_end > seq_size)
    {
      fflush(stdout);
      fprintf(stderr, "twoBitReadSeqFrag in %s end (%d) >= seqSize (%d)\n", s->name, frag_end, seq_size);
      exit(255);
    }
    if (frag_end - flanking_start < 1)
    {
      fflush(stdout);
      fprintf(stderr, "twoBitReadSeqFrag in %s start (%d) >= end (%d)\n", s->name, flanking_start, frag_end);
      exit(255);
    }
    s->win_start = flanking_start;
    s->win_end = frag_end;
    s->lower_flank_len = lower_flank_len;
    s->lower_flag = lower_flag;
    s->upper_flag = upper_flag;
    total += (uint64_t) (frag_end - flanking_start);
  }

  /* ---- one allocation per array, decode, cores ------------------------------------------------- */
  struct sequenceLibrary *lib = (struct sequenceLibrary *) calloc(1, sizeof(struct sequenceLibrary));
  char *sequence = (char *) malloc(total + 1);
  char **idents = (char **) malloc(((size_t) n + 1) * sizeof(char *));
  uint64_t *boundaries = (uint64_t *) malloc(((size_t) n + 1) * sizeof(uint64_t));
  uint64_t *offsets = (uint64_t *) malloc(((size_t) n + 1) * sizeof(uint64_t));
  struct coreAlignment *cores = (struct coreAlignment *) calloc((size_t) (n > 0 ? n : 1), sizeof(struct coreAlignment));
  if (!lib || !sequence || !idents || !boundaries || !offsets || !cores)
  {
    fprintf(stderr, "Could not allocated more space for the sequence array\n");
    exit(1);
  }

  /* device source: every window's packed bytes, copied from the 16-byte block its first base lies in */
  free(g_source.blob);
  free(g_source.windows);
  free(g_source.n_start);
  free(g_source.n_len);
  memset(&g_source, 0, sizeof(g_source));
  uint64_t blob_bytes = 0;
  for (int i = 0; i < n; ++i)
  {
    const uint64_t b0 = ((uint64_t) rg[i].win_start >> 2) & ~15ull, b1 = ((uint64_t) rg[i].win_end + 3) >> 2;
    blob_bytes += ((b1 - b0) + 15) & ~15ull;
  }
  uint8_t *blob = NULL;
  if (posix_memalign((void **) &blob, 64, (size_t) blob_bytes + 64) != 0)
    blob = NULL;
  ramx_window_t *windows = (ramx_window_t *) malloc((size_t) (n > 0 ? n : 1) * sizeof(ramx_window_t));
  if (!blob || !windows)
  {
    fprintf(stderr, "Could not allocated more space for the packed windows\n");
    exit(1);
  }
  memset(blob + blob_bytes, 0, 64);

  uint64_t pos = 0, bpos = 0;
  for (int i = 0; i < n; ++i)
  {
    const struct range *s = &rg[i];
    const struct record *rec = &tb.rec[s->record];
    const uint64_t len = (uint64_t) (s->win_end - s->win_start);
    decode_window(&tb, rec, s->win_start, s->win_end, sequence + pos);
    /* N blocks overwrite (kentsrc/twoBitNew.c:597-613); everything that is not ACGT is 99 */
    /* the blocks of a record are sorted and disjoint (the reference relies on it too:
     * findGreatestLowerBound, kentsrc/twoBitNew.c:597): start at the last block that begins at or
     * before the window, stop at the first one that begins behind it */
    uint32_t b_lo = 0, b_hi = rec->n_blocks;
    while (b_lo < b_hi)
    {
      const uint32_t mid = b_lo + (b_hi - b_lo) / 2;
      if ((long) rd32(&tb, rec->nstart_off + 4ull * mid) <= s->win_start)
        b_lo = mid + 1;
      else
        b_hi = mid;
    }
    for (uint32_t b = b_lo > 0 ? b_lo - 1 : 0; b < rec->n_blocks; ++b)
    {
      const long ns = (long) rd32(&tb, rec->nstart_off + 4ull * b);
      if (ns >= s->win_end)
        break;
      const long ne = ns + (long) rd32(&tb, rec->nsize_off + 4ull * b);
      const long a = ns > s->win_start ? ns : s->win_start, e = ne < s->win_end ? ne : s->win_end;
      if (a < e)
      {
        memset(sequence + pos + (a - s->win_start), 99, (size_t) (e - a));
        nrange_push(pos + (uint64_t) (a - s->win_start), (uint64_t) (e - a));
      }
    }
    {
      const uint64_t b0 = ((uint64_t) s->win_start >> 2) & ~15ull, b1 = ((uint64_t) s->win_end + 3) >> 2;
      memcpy(blob + bpos, tb.map + rec->dna_off + b0, (size_t) (b1 - b0));
      const uint64_t padded = ((b1 - b0) + 15) & ~15ull;
      memset(blob + bpos + (b1 - b0), 0, (size_t) (padded - (b1 - b0)));
      windows[i].dst = pos;
      windows[i].len = len;
      windows[i].src_byte = bpos;
      windows[i].src_base = (uint64_t) s->win_start - 4 * b0;
      bpos += padded;
    }

    idents[i] = strdup(s->name);
    boundaries[i] = pos + len;
    offsets[i] = (uint64_t) s->win_start;

    struct coreAlignment *c = &cores[i];
    c->next = (i + 1 < n) ? &cores[i + 1] : NULL;
    c->seqIdx = i;
    c->lowerSeqBound = pos;
    c->upperSeqBound = pos + len - 1;
    c->lowerSeqBoundFlag = (enum CoreBoundFlag) s->lower_flag;
    c->upperSeqBoundFlag = (enum CoreBoundFlag) s->upper_flag;
    c->leftExtendable = s->left_flag ? 1 : 0;
    c->rightExtendable = s->right_flag ? 1 : 0;
    c->leftExtensionLen = 0;
    c->rightExtensionLen = 0;
    c->score = 0;
    if (s->reverse)   /* sequence.c:886-891 */
    {
      c->orient = 1;
      c->rightSeqPos = pos + (uint64_t) s->lower_flank_len;
      c->leftSeqPos = c->rightSeqPos + (uint64_t) (s->end - s->start) - 1;
    }
    else              /* sequence.c:892-897 */
    {
      c->orient = 0;
      c->leftSeqPos = pos + (uint64_t) s->lower_flank_len;
      c->rightSeqPos = c->leftSeqPos + (uint64_t) (s->end - s->start) - 1;
    }
    pos += len;
  }
  idents[n] = NULL;
  boundaries[n] = 0;
  offsets[n] = 0;

  if (n > 0)
    *core_align = cores;
  *num_cores = n;
  lib->sequence = sequence;
  lib->identifiers = idents;
  lib->boundaries = boundaries;
  lib->offsets = offsets;
  lib->length = total;
  lib->count = n;

  g_last.lib = lib;
  g_last.cores = cores;
  g_last.ranges = rg;
  g_last.n = n;
  g_source.sequence = sequence;
  g_source.total = total;
  g_source.blob = blob;
  g_source.blob_bytes = blob_bytes + 64;
  g_source.windows = windows;
  g_source.n_windows = n;

  munmap((void *) tb.map, tb.bytes);
  free(tb.rec);
  if (getenv("RAMX_TIMING"))
  {
    struct timespec t_end;
    clock_gettime(CLOCK_MONOTONIC, &t_end);
    fprintf(stderr, "[ramx timing] loader %.3f s (%d cores, %llu bases)\n",
            (double) (t_end.tv_sec - t_begin.tv_sec) + 1e-9 * (double) (t_end.tv_nsec - t_begin.tv_nsec), n,
            (unsigned long long) total);
  }
  /* rg[].name point into the ranges buffer: both stay allocated until exit, like the reference's list */
  return lib;
}

void
ramx_loader_release_last(void)
{
  if (g_last.lib)
  {
    for (int i = 0; i < g_last.n; ++i)
      free(g_last.lib->identifiers[i]);
    free(g_last.lib->identifiers);
    free(g_last.lib->sequence);
    free(g_last.lib->boundaries);
    free(g_last.lib->offsets);
    free(g_last.lib);
  }
  free(g_last.cores);
  free(g_last.ranges);
  free(g_last.ranges_text);
  memset(&g_last, 0, sizeof(g_last));
  free(g_source.blob);
  free(g_source.windows);
  free(g_source.n_start);
  free(g_source.n_len);
  memset(&g_source, 0, sizeof(g_source));
}

/* ---------------------------------------------------------------------------------------------
 * main() also allocates the reference's DP scratch, score[2][MAXN][2B+1][2] as 2 * MAXN * (2B + 2)
 * mallocs (bnw_extend.c:87-135; MAXN = 10000 whatever the family size). The GPU path keeps its rows
 * in registers and never reads it: hand out a token instead.
 * ------------------------------------------------------------------------------------------- */
int ****
allocate_score(int N, int bandwidth)
{
  static int *token[1];
  (void) N;
  (void) bandwidth;
  return (int ****) token;
}

/* referenced by the reference's own (weakened, unused) extend_alignment() body in ram_extend.o */
int
compute_nw_row(int direction, int row_idx, int n, char cons_base, struct coreAlignment *coreAlign, int ****score,
               uint64_t lower_seq_bound, uint64_t upper_seq_bound, char *sequence, int *max_score_sequence_idx,
               struct scoringSystem *scoreParams, int bandwidth, int L, int VERBOSE, char *pathString)
{
  (void) direction; (void) row_idx; (void) n; (void) cons_base; (void) coreAlign; (void) score;
  (void) lower_seq_bound; (void) upper_seq_bound; (void) sequence; (void) max_score_sequence_idx;
  (void) scoreParams; (void) bandwidth; (void) L; (void) VERBOSE; (void) pathString;
  printf("RAMExtend (CUDA build): compute_nw_row() is not part of this build; the band rows are computed on the GPU\n");
  exit(1);
}
