/* rr_emu.cpp -- TEST INFRASTRUCTURE ONLY: host build of the step-kernel text with the 32 lanes run as fibers.
 *
 * There is no GPU in the development container, so the device code in brax_rodent_run_b200/csrc/rr_kernels.inl
 * is written against a handful of warp primitives (__shfl_sync, __shfl_xor_sync, __ballot_sync, __syncwarp,
 * __ldg).  This file supplies those primitives on the host: the 32 lanes of an environment's warp are ucontext
 * fibers scheduled round-robin, every primitive is a yield point, and exchanged values go through a
 * double-buffered slot array.  It exports the same C ABI as librr_b200.so (include/rr_b200.h) with "device"
 * pointers being host pointers, so tests/ can drive the identical host logic + kernel text on the CPU and compare
 * it with the oracle.  It is NOT a CPU fallback: the product package never loads it (the Python layer refuses
 * to run without the CUDA library unless a test passes the emulator handle explicitly).
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <vector>

/* ---- warp emulation ---------------------------------------------------------------------------------- */
namespace emu {
static const int NL = 32;
static ucontext_t g_sched, g_fib[NL];
static int g_lane = 0;
static bool g_done[NL];
static int g_par[NL];
static uint32_t g_slot[2][NL];
static int g_tag[2][NL];
static std::vector<char> g_stacks;
static void (*g_body)(int lane, void *arg);
static void *g_arg;

static inline void yield() { swapcontext(&g_fib[g_lane], &g_sched); }

static inline uint32_t exchange(uint32_t v, int src, int tag) {
  int l = g_lane, p = g_par[l];
  g_slot[p][l] = v;
  g_tag[p][l] = tag;
  g_par[l] ^= 1;
  yield();
  if (g_tag[p][src & 31] != tag) { fprintf(stderr, "emu: lanes diverged at a warp collective (tag %d vs %d)\n", tag, g_tag[p][src & 31]); abort(); }
  return g_slot[p][src & 31];
}

static void trampoline() {
  int l = g_lane;
  g_body(l, g_arg);
  g_done[l] = true;
  swapcontext(&g_fib[l], &g_sched);
}

static void run_warp(void (*body)(int, void *), void *arg) {
  const size_t STK = 512 * 1024;
  if (g_stacks.size() < STK * NL) g_stacks.resize(STK * NL);
  g_body = body;
  g_arg = arg;
  for (int l = 0; l < NL; l++) {
    getcontext(&g_fib[l]);
    g_fib[l].uc_stack.ss_sp = g_stacks.data() + STK * l;
    g_fib[l].uc_stack.ss_size = STK;
    g_fib[l].uc_link = &g_sched;
    makecontext(&g_fib[l], trampoline, 0);
    g_done[l] = false;
    g_par[l] = 0;
  }
  /* RR_EMU_LANE_ORDER=reverse|odd-first runs the lanes of every scheduling round in another order: a hazard between lanes
   * that is not separated by a warp primitive (the class of bug racecheck finds) then changes the result */
  const char *ord = getenv("RR_EMU_LANE_ORDER");
  const int mode = !ord ? 0 : (ord[0] == 'r' ? 1 : 2);
  for (;;) {
    int alive = 0;
    for (int k = 0; k < NL; k++) {
      const int l = mode == 0 ? k : (mode == 1 ? NL - 1 - k : (k < NL / 2 ? 2 * k + 1 : 2 * (k - NL / 2)));
      if (g_done[l]) continue;
      g_lane = l;
      swapcontext(&g_sched, &g_fib[l]);
      alive++;
    }
    if (!alive) break;
  }
}
}  // namespace emu

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline float __shfl_sync(unsigned, float v, int src) { return u2f(emu::exchange(f2u(v), src, 1)); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)emu::exchange((uint32_t)v, src, 2); }
static inline float __shfl_xor_sync(unsigned, float v, int mask) { return u2f(emu::exchange(f2u(v), emu::g_lane ^ mask, 3)); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
  int l = emu::g_lane, p = emu::g_par[l];
  emu::g_slot[p][l] = pred ? 1u : 0u;
  emu::g_tag[p][l] = 4;
  emu::g_par[l] ^= 1;
  emu::yield();
  unsigned r = 0;
  for (int k = 0; k < 32; k++) {
    if (emu::g_tag[p][k] != 4) { fprintf(stderr, "emu: lanes diverged at ballot\n"); abort(); }
    r |= emu::g_slot[p][k] << k;
  }
  return r;
}
static inline void __syncwarp() { emu::exchange(0, emu::g_lane, 5); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }

#define RR_DEV static inline
#define RR_HOSTDEV static inline
#define RR_DEV_MEMBER inline
#define RR_DEV_NOINLINE static
#define RR_LDG(p) (*(p))
#define RR_CLOCK() 0LL
#define RR_CTA_SYNC() ((void)0)
typedef const char *rr_emu_saddr;
#define RR_SADDR_T rr_emu_saddr
#define RR_SADDR(p) ((rr_emu_saddr)(p))
#define RR_SLOAD(a) (*(const float *)(a))
#define RR_RCP(x) (1.f / (x))
#define RR_SQRT(x) sqrtf(x)
#define RR_SINCOS(x, s, c) sincosf(x, s, c)
#define RR_SOLVE_UPDATE(x, m, bit, a, xi) do { if ((m) & (bit)) (x) -= RR_SLOAD(a) * (xi); } while (0)

#include "../../brax_rodent_run_b200/csrc/rr_kernels.inl"

/* ---- host backend for the shared C-ABI body ------------------------------------------------------------- */
static const char *rrb_error() { return "emulator backend error"; }
static int rrb_set_device(int) { return 0; }
static int rrb_malloc(void **p, size_t bytes) { *p = calloc(bytes ? bytes : 4, 1); return *p ? 0 : 1; }
static void rrb_free(void *p) { free(p); }
static int rrb_h2d(void *dst, const void *src, size_t bytes, void *) { memcpy(dst, src, bytes); return 0; }
static int rrb_d2h(void *dst, const void *src, size_t bytes, void *) { memcpy(dst, src, bytes); return 0; }
static void *rrb_host_devptr(void *) { return nullptr; }
static int rrb_sync(void *) { return 0; }
static int rrb_num_slots() { return 1; }
static int rrb_fp32_peak(double *, void *) { return 1; } /* no device: the call fails with RR_ECUDA */
static int rrb_geometry(const RRModelDev &, int B, int *ctas, int *wpb) { *ctas = B; *wpb = 1; return 0; }

struct EmuJob { const RRModelDev *m; const RRStepArgs *a; int env; float *sm; const int32_t *ti; const float *tf; };
template <int NS>
static void emu_lane(int lane, void *arg) {
  EmuJob *j = (EmuJob *)arg;
  rr::env_run<NS>(*j->m, *j->a, j->env, 0, j->sm, j->ti, j->tf, lane);
}
static int rrb_launch_step(const RRModelDev &m, const RRStepArgs &a, void *) {
  std::vector<float> sm((size_t)m.sm.total + 16);
  for (int slot = 0; slot < a.B; slot++) {
    int env = a.env_order ? a.env_order[slot] : slot;
    if (env < 0) continue;
    /* poison shared memory so that reads of unwritten slots show up as NaN */
    for (auto &x : sm) x = NAN;
    /* "shared memory" copies of the tables, as the CUDA kernel stages them */
    std::vector<int32_t> ti(m.ibuf, m.ibuf + m.ni);
    std::vector<float> tf(m.fbuf, m.fbuf + m.nf);
    EmuJob j{&m, &a, env, sm.data(), ti.data(), tf.data()};
    emu::run_warp(m.nv <= 96 ? emu_lane<3> : emu_lane<5>, &j);
  }
  return 0;
}
static int rrb_launch_gae(const float *rewards, const float *values, const float *bootstrap, const float *termination,
                          const float *truncation, int T, int B, float discount, float lambda_, float *vs, float *adv, void *) {
  for (int b = 0; b < B; b++) {
    float acc = 0.f, v_next = bootstrap[b], vs_next = bootstrap[b];
    for (int t = T - 1; t >= 0; t--) {
      size_t i = (size_t)t * B + b;
      float mask = 1.f - truncation[i], term = termination[i], v = values[i], r = rewards[i];
      float delta = (r + discount * (1.f - term) * v_next - v) * mask;
      acc = delta + discount * (1.f - term) * mask * lambda_ * acc;
      float vs_t = acc + v;
      adv[i] = (r + discount * (1.f - term) * vs_next - v) * mask;
      vs[i] = vs_t;
      v_next = v;
      vs_next = vs_t;
    }
  }
  return 0;
}

#define RR_PPO_HD static inline
#include "../../brax_rodent_run_b200/csrc/rr_ppo_loss.h"
static int rrb_ppo_blocks(int n) { return (n + 127) / 128; }
static int rrb_ppo_blocks_b(int n) { return rrb_ppo_blocks(n); }
static int rrb_launch_ppo_loss(const RRPpoLossArgs &a, void *) {
  /* same block decomposition and per-block sums as the CUDA kernels (sequential inside a block) */
  for (int blk = 0; blk < rrb_ppo_blocks(a.B); blk++) {
    double s1 = 0.0, s2 = 0.0;
    for (int b = blk * 128; b < a.B && b < (blk + 1) * 128; b++) {
      double t1, t2;
      rr_ppo_stage_a(a, b, t1, t2);
      s1 += t1; s2 += t2;
    }
    a.adv_partial[2 * blk] = s1; a.adv_partial[2 * blk + 1] = s2;
  }
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < a.nblkA; k++) { s1 += a.adv_partial[2 * k]; s2 += a.adv_partial[2 * k + 1]; }
  const double n = (double)a.T * (double)a.B, mean = s1 / n;
  double var = s2 / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const size_t N = (size_t)a.T * a.B;
  for (int blk = 0; blk < rrb_ppo_blocks((int)N); blk++) {
    float acc[3] = {0.f, 0.f, 0.f};
    for (size_t i = (size_t)blk * 128; i < N && i < (size_t)(blk + 1) * 128; i++) {
      float pol, val, ent;
      rr_ppo_stage_b(a, i, (float)mean, (float)sqrt(var), pol, val, ent);
      acc[0] += pol; acc[1] += val; acc[2] += ent;
    }
    for (int q = 0; q < 3; q++) a.loss_partial[3 * blk + q] = acc[q];
  }
  return 0;
}

/* Adam and the minibatch gather: the same per-element arithmetic in plain loops */
#define RR_MISC_HD static inline
#include "../../brax_rodent_run_b200/csrc/rr_learner_misc.h"
static int rrb_adam_step(float *p, float *g, const float *partials, int nsplit, float *m, float *v, float *step, long long n, float lr,
                         float b1, float b2, float eps, void *) {
  const float t = step[0] + 1.f;
  for (long long i = 0; i < n; i++) {
    if (partials) {
      float gi = partials[i];
      for (int s = 1; s < nsplit; s++) gi += partials[(size_t)s * n + i];
      g[i] = gi;
    }
    rr_adam_element(p[i], g[i], m[i], v[i], t, lr, b1, b2, eps);
  }
  step[0] = t;
  return 0;
}
static int rrb_policy_act(const rr_policy_args &a, void *) { rr_policy_reference(a); return 0; }
static int rrb_gather_rows(const RRGatherArgs &a, int, void *) {
  for (int it = 0; it < a.count; it++) {
    const rr_gather_item &g = a.item[it];
    for (int t = 0; t < g.outer; t++)
      for (int j = 0; j < a.rows; j++)
        memcpy(g.dst + ((size_t)t * a.rows + j) * g.dst_pitch, g.src + ((size_t)t * g.src_rows + a.idx[j]) * g.inner, sizeof(float) * g.inner);
  }
  return 0;
}

/* the learner's grouped GEMM: plain loops with the contract of rr_tc_problem ("device" pointers are host pointers) */
#define RR_TC_HD static inline
#include "../../brax_rodent_run_b200/csrc/rr_tc_gemm.h"
static int rrb_tc_smem_max() { return RR_TC_STAGES * (RR_TC_BM + 128) * RR_TC_BK * 4; }
static void rrb_tc_encode(RRTcRecord &) {} /* no TMA on the host */
static int rrb_tc_launch(const RRTcRecord *recs, int count, int, int, void *) {
  for (int i = 0; i < count; i++) rr_tc_reference(recs[i].p);
  return 0;
}

#include "../../brax_rodent_run_b200/csrc/rr_api_impl.inl"
