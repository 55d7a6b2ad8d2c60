/* ramx_boundary.h -- what csrc/ramx_boundary.c offers to the batch front-end (csrc/ramx_batch.c). */
#ifndef RAMX_BOUNDARY_H
#define RAMX_BOUNDARY_H
#include <stdint.h>
#include "ramcu.h"

struct sequenceLibrary;

/* one extend_alignment() call, flattened: return 1 with cons / ext_len / score_add / res filled, or 0 to
 * let the boundary run the call itself */
typedef int (*ramx_extend_hook_fn)(int direction, int n_cores, const ramcu_core_t *cores,
                                   const ramcu_scoring_t *scoring, const ramcu_params_t *params,
                                   struct sequenceLibrary *seqLib, uint8_t *cons, int32_t *ext_len,
                                   int32_t *score_add, ramcu_result_t *res);
extern ramx_extend_hook_fn ramx_extend_hook;

/* the boundary's CUDA context (created on first use; RAMCU_DEVICE selects the GPU) */
ramcu_ctx *ramx_boundary_context(void);
void ramx_boundary_forget_library(void);

#endif
