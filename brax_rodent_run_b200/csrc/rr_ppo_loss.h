/* rr_ppo_loss.h -- the PPO minibatch loss and its gradient w.r.t. the network outputs, fused.
 *
 * Restates brax ppo.losses.compute_ppo_loss (called inside ppo.train, brax_rodent_run_ppo.py:200):
 * NormalTanhDistribution log-prob of the taken raw action, compute_gae, advantage normalisation, clipped surrogate, value
 * loss (0.5 * 0.5 * mse against vs), entropy of the tanh-normal estimated with one noise sample, and -- instead of
 * autograd over ~130 element-wise / reduction launches -- the analytic gradient w.r.t. logits and baseline.  Two stages:
 *   A  one environment column: GAE scan over T, sum / sum of squares of the advantages
 *   B  one (t, env) element: log-prob, losses and gradients
 * The same text runs as two CUDA kernels (rr_api.cu) and as host loops in the test emulator (tests/emu/rr_emu.cpp).
 */
#ifndef RR_PPO_LOSS_H_
#define RR_PPO_LOSS_H_

#include <math.h>
#include <stdint.h>

#ifndef RR_PPO_HD
#define RR_PPO_HD __host__ __device__ inline
#endif

struct RRPpoLossArgs {
  int T, B, A;
  const float *logits;       /* [T, B, 2 A]  (loc | pre-softplus scale) */
  const float *baseline;     /* [T, B] */
  const float *bootstrap;    /* [B] */
  const float *raw_action;   /* [T, B, A] */
  const float *old_log_prob; /* [T, B] */
  const float *reward, *discount, *truncation; /* [T, B] */
  const float *noise;        /* [T, B, A] entropy sample */
  float reward_scaling, gamma, lambda_, clip_eps, entropy_cost;
  int normalize_advantage;
  float *lp, *adv, *vs;      /* [T, B] scratch */
  double *adv_partial;       /* [2 nblkA] sum, sum of squares per stage-A block */
  float *loss_partial;       /* [3 nblkB] policy, value, entropy sums per stage-B block */
  float *grad_logits;        /* [T, B, 2 A] */
  float *grad_baseline;      /* [T, B] */
  int nblkA;
};

#define RR_PPO_LOG2 0.6931471805599453f
#define RR_PPO_HALF_LOG_2PI 0.9189385332046727f

RR_PPO_HD float rr_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); } /* torch.nn.functional.softplus */
RR_PPO_HD float rr_log_det_tanh(float x) { return 2.f * (RR_PPO_LOG2 - x - rr_softplus(-2.f * x)); }

/* stage A for column b: the GAE scan (sequential over T, nothing transcendental); returns (sum adv, sum adv^2) */
RR_PPO_HD void rr_ppo_stage_a(const RRPpoLossArgs &a, int b, double &s1, double &s2) {
  const int T = a.T, B = a.B;
  float acc = 0.f, v_next = a.bootstrap[b], vs_next = a.bootstrap[b];
  s1 = 0.0; s2 = 0.0;
  for (int t = T - 1; t >= 0; t--) {
    const size_t i = (size_t)t * B + b;
    const float trunc = a.truncation[i], term = (1.f - a.discount[i]) * (1.f - trunc), mask = 1.f - trunc;
    const float v = a.baseline[i], r = a.reward[i] * a.reward_scaling;
    const float delta = (r + a.gamma * (1.f - term) * v_next - v) * mask;
    acc = delta + a.gamma * (1.f - term) * mask * a.lambda_ * acc;
    const float vs_t = acc + v, adv = (r + a.gamma * (1.f - term) * vs_next - v) * mask;
    a.adv[i] = adv;
    a.vs[i] = vs_t;
    v_next = v;
    vs_next = vs_t;
    s1 += (double)adv;
    s2 += (double)adv * (double)adv;
  }
}

/* stage B for element i = t B + b; mean / std of the advantages from the stage-A partials; returns the three loss terms */
RR_PPO_HD void rr_ppo_stage_b(const RRPpoLossArgs &a, size_t i, float mean, float std_, float &pol, float &val, float &ent) {
  const int A = a.A;
  const float invN = 1.f / ((float)a.T * (float)a.B);
  const float adv = a.normalize_advantage ? (a.adv[i] - mean) / (std_ + 1e-8f) : a.adv[i];
  const float *lg = a.logits + i * 2 * A, *raw = a.raw_action + i * A, *nz = a.noise + i * A;
  float lp = 0.f; /* log-prob of the taken action under the current policy (first pass over the actions) */
  for (int k = 0; k < A; k++) {
    const float scale = rr_softplus(lg[A + k]) + 1e-3f, z = (raw[k] - lg[k]) / scale;
    lp += -0.5f * z * z - logf(scale) - RR_PPO_HALF_LOG_2PI - rr_log_det_tanh(raw[k]);
  }
  a.lp[i] = lp;
  const float rho = expf(lp - a.old_log_prob[i]);
  const float lo = 1.f - a.clip_eps, hi = 1.f + a.clip_eps;
  const float s1 = rho * adv, s2 = fminf(fmaxf(rho, lo), hi) * adv;
  pol = -fminf(s1, s2);
  const float g = (s1 <= s2) ? -adv * rho * invN : 0.f; /* d total / d log-prob (a tie outside the clip range has adv = 0) */
  const float d = a.vs[i] - a.baseline[i];
  val = 0.25f * d * d;
  a.grad_baseline[i] = -0.5f * d * invN;
  float *gl = a.grad_logits + i * 2 * A;
  const float ce = -a.entropy_cost * invN; /* d total / d entropy element */
  float e = 0.f;
  for (int k = 0; k < A; k++) {
    const float pre = lg[A + k], scale = rr_softplus(pre) + 1e-3f, is = 1.f / scale;
    const float sig = pre > 20.f ? 1.f : 1.f / (1.f + expf(-pre));
    const float z = (raw[k] - lg[k]) * is;
    const float raw_e = lg[k] + scale * nz[k], th = tanhf(raw_e);
    e += 0.5f + RR_PPO_HALF_LOG_2PI + logf(scale) + rr_log_det_tanh(raw_e);
    gl[k] = g * z * is + ce * (-2.f * th);
    gl[A + k] = (g * (z * z - 1.f) * is + ce * (is - 2.f * th * nz[k])) * sig;
  }
  ent = e;
}

#endif /* RR_PPO_LOSS_H_ */
