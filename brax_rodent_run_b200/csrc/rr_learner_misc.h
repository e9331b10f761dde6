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

/* ---- policy inference (rr_policy_act): the per-element pieces shared by the CUDA kernel and the host loops ---- */
#define RR_POLICY_HIDDEN 32
RR_MISC_HD float rr_policy_silu(float x) { return x / (1.f + expf(-x)); }
RR_MISC_HD float rr_policy_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); } /* torch.nn.functional.softplus */
/* NormalTanhDistribution for one action dimension: returns the log-prob term; deterministic when eps is null */
RR_MISC_HD float rr_policy_sample(float loc, float pre, const float *eps, float &action, float &raw) {
  if (!eps) { raw = loc; action = tanhf(loc); return 0.f; }
  const float scale = rr_policy_softplus(pre) + 1e-3f;
  raw = loc + scale * eps[0];
  action = tanhf(raw);
  const float z = (raw - loc) / scale;
  const float log_det = 2.f * (0.6931471805599453f - raw - rr_policy_softplus(-2.f * raw));
  return -0.5f * z * z - logf(scale) - 0.9189385332046727f - log_det;
}
/* plain loops: the emulator backend and the statement of the contract */
static inline void rr_policy_reference(const rr_policy_args &a) {
  const int H = RR_POLICY_HIDDEN;
  for (int r = 0; r < a.B; r++) {
    float h[64], t[64];
    for (int j = 0; j < H; j++) {
      float acc = a.b[0][j];
      for (int k = 0; k < a.obs_dim; k++) {
        const float x = a.obs[(size_t)r * a.obs_dim + k];
        acc += (a.mean ? (x - a.mean[k]) / a.std[k] : x) * a.w[0][(size_t)j * a.in0 + k];
      }
      h[j] = rr_policy_silu(acc);
    }
    for (int l = 1; l < a.nlayers; l++) {
      const int nout = l == a.nlayers - 1 ? 2 * a.A : H;
      for (int j = 0; j < nout; j++) {
        float acc = a.b[l][j];
        for (int k = 0; k < H; k++) acc += h[k] * a.w[l][j * H + k];
        t[j] = l == a.nlayers - 1 ? acc : rr_policy_silu(acc);
      }
      for (int j = 0; j < nout; j++) h[j] = t[j];
    }
    float lp = 0.f;
    for (int k = 0; k < a.A; k++)
      lp += rr_policy_sample(h[k], h[a.A + k], a.eps ? a.eps + (size_t)r * a.A + k : nullptr, a.action[(size_t)r * a.A + k],
                             a.raw_action[(size_t)r * a.A + k]);
    a.log_prob[r] = lp;
  }
}
