/* ramx_loader.h -- what ramx_loader.c (the product's loadSequenceSubsetMinimal) keeps for the boundary:
 * the windows it extracted, as the .2bit file's own packed bytes, and the N runs inside them. */
#ifndef RAMX_LOADER_H
#define RAMX_LOADER_H
#include <stdint.h>

/* same layout as ramcu_window_t (include/ramcu.h) */
typedef struct
{
  uint64_t dst;       /* first library position of the window */
  uint64_t len;       /* bases */
  uint64_t src_byte;  /* byte offset inside `blob` of the 16-byte block the window's first base lies in */
  uint64_t src_base;  /* index of the window's first base counted from that byte */
} ramx_window_t;

struct ramx_twobit_source
{
  const char *sequence;   /* the seqLib->sequence these windows describe */
  uint64_t total;         /* == seqLib->length */
  uint8_t *blob;          /* packed DNA, 4 bases per byte, T0 C1 A2 G3 (kentsrc/twoBitNew.c:571-582) */
  uint64_t blob_bytes;
  ramx_window_t *windows;
  int n_windows;
  uint64_t *n_start, *n_len;   /* N runs in library coordinates */
  int n_nranges, cap_nranges;
};

/* the source of the library (sequence, length) if the product loader built it, else NULL */
const struct ramx_twobit_source *ramx_loader_source(const char *sequence, uint64_t length);

/* free everything the last loadSequenceSubsetMinimal() call allocated (batch front-end only: the
 * reference's main() keeps using the library until it returns) */
void ramx_loader_release_last(void);

#endif
