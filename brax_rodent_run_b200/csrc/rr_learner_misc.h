/* rr_learner_misc.h -- the small kernels around the learner's GEMMs: Adam on the flat parameter buffer and the minibatch gather.
 * The per-element functions are shared by the CUDA kernels (rr_api.cu) and the emulator backend (tests/emu). */
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/rr_b200.h"

#define RR_GATHER_MAX 8
struct RRGatherArgs {
  rr_gather_item item[RR_GATHER_MAX];
  int32_t block_start[RR_GATHER_MAX + 1]; /* one block per (item, t, j) row */
  int32_t count, rows;
  const int64_t *idx;
};

/* torch.optim.Adam (single tensor, no weight decay, no amsgrad), step = the 1-based count of this step */
RR_MISC_HD void rr_adam_element(float &p, float g, float &m, float &v, float step, float lr, float b1, float b2, float eps) {
  m = m + (g - m) * (1.f - b1);           /* exp_avg.lerp_(grad, 1 - beta1) */
  v = v * b2 + g * g * (1.f - b2);        /* exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2) */
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float denom = sqrtf(v) / sqrtf(bc2) + eps;
  p -= (lr / bc1) * (m / denom);
}
