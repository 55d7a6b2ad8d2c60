/* ramx_loader_flat.c -- flat-array entry point over the product loader, for the tests
 * (host/_refobj/libramxload.so): same protocol as oracle/ref_shim.c:ramref_load, so that the two
 * loaders can be diffed on the same files. Call with seq_out == NULL to get the sizes. */
#include <stdint.h>
#include <string.h>

#include "common.h"
#include "sequence.h"
#include "ramx_loader.h"

typedef struct
{
  int64_t left_seq_pos, right_seq_pos, lower_seq_bound, upper_seq_bound;
  int8_t orient, left_extendable, right_extendable, pad_[5];
} flat_core_t;

int
ramx_load_flat(const char *twobit, const char *ranges, int max_flanking_bp, int *n_out, uint64_t *seq_len_out,
               flat_core_t *cores_out, char *seq_out, uint64_t *offsets_out, uint64_t *boundaries_out,
               int32_t *flags_out, uint64_t *n_windows_out, uint64_t *n_nranges_out)
{
  struct coreAlignment *ca = NULL, *c;
  int n = 0, i;
  struct sequenceLibrary *lib = loadSequenceSubsetMinimal((char *) twobit, (char *) ranges, &ca, &n, max_flanking_bp);
  *n_out = n;
  *seq_len_out = lib->length;
  if (seq_out)
  {
    const struct ramx_twobit_source *src = ramx_loader_source(lib->sequence, lib->length);
    memcpy(seq_out, lib->sequence, lib->length);
    for (c = ca, i = 0; c != NULL; c = c->next, i++)
    {
      memset(&cores_out[i], 0, sizeof(flat_core_t));
      cores_out[i].left_seq_pos = (int64_t) c->leftSeqPos;
      cores_out[i].right_seq_pos = (int64_t) c->rightSeqPos;
      cores_out[i].lower_seq_bound = (int64_t) c->lowerSeqBound;
      cores_out[i].upper_seq_bound = (int64_t) c->upperSeqBound;
      cores_out[i].orient = c->orient;
      cores_out[i].left_extendable = c->leftExtendable;
      cores_out[i].right_extendable = c->rightExtendable;
      if (offsets_out) offsets_out[i] = lib->offsets[i];
      if (boundaries_out) boundaries_out[i] = lib->boundaries[i];
      if (flags_out)
      {
        flags_out[3 * i] = (int32_t) c->lowerSeqBoundFlag;
        flags_out[3 * i + 1] = (int32_t) c->upperSeqBoundFlag;
        flags_out[3 * i + 2] = c->seqIdx;
      }
    }
    if (n_windows_out) *n_windows_out = src ? (uint64_t) src->n_windows : 0;
    if (n_nranges_out) *n_nranges_out = src ? (uint64_t) src->n_nranges : 0;
  }
  return n;
}

/* the packed windows the loader keeps for the device: copies them out (tests decode them with numpy) */
int
ramx_load_source(const char *seq_key_unused, uint8_t *blob_out, uint64_t *blob_bytes, uint64_t *windows_out,
                 uint64_t *n_start_out, uint64_t *n_len_out)
{
  (void) seq_key_unused;
  extern const struct ramx_twobit_source *ramx_loader_last_source(void);
  const struct ramx_twobit_source *src = ramx_loader_last_source();
  if (!src)
    return -1;
  *blob_bytes = src->blob_bytes;
  if (blob_out)
  {
    memcpy(blob_out, src->blob, src->blob_bytes);
    memcpy(windows_out, src->windows, (size_t) src->n_windows * sizeof(ramx_window_t));
    memcpy(n_start_out, src->n_start, (size_t) src->n_nranges * sizeof(uint64_t));
    memcpy(n_len_out, src->n_len, (size_t) src->n_nranges * sizeof(uint64_t));
  }
  return src->n_windows;
}
