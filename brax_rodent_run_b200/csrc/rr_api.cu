/* rr_api.cu -- CUDA (sm_100a) backend of the C ABI in include/rr_b200.h.
 *
 * Build (see brax_rodent_run_b200/build.py):
 *   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC -o librr_b200.so rr_api.cu
 *
 * Launch geometry: one warp per environment, RR_WPB warps per CTA; each warp owns `sm.total` floats of dynamic
 * shared memory holding the whole per-environment working set of a physics substep (about 33 KB for
 * rodent_0.xml), so HBM is touched only to load the 1 KB state + action and to store state + observation.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <math_constants.h>
#include <atomic>

/* production kernels: debug-dump / clock64 hooks compiled out */
#define RR_NS rr
#define RR_WITH_DEBUG 0
#include "rr_kernels.inl"
#undef RR_NS
#undef RR_WITH_DEBUG
/* instrumented kernels (parity tests' intermediates, per-phase cycle counts) */
#define RR_NS rr_dbg
#define RR_WITH_DEBUG 1
#include "rr_kernels.inl"
#undef RR_NS
#undef RR_WITH_DEBUG

/* Environments (warps) per CTA.  rodent_0 fits 10 in shared memory; bounding the launch at 10 warps (not 12) also lets
 * ptxas keep a few more values in registers (measured +3.8 %). */
#ifndef RR_MAX_WPB
#define RR_MAX_WPB 10
#endif
#define RR_SMEM_MAX 232448 /* 227 KB opt-in dynamic shared memory per CTA on sm_100 */

/* Persistent CTAs: the model tables (about 29 KB for rodent_0) are staged into shared memory once per CTA, then each
 * warp loops over its share of the environments. */
/* Models with more than 96 dofs (NS = 5 register slots per nv-vector, e.g. rodent_pair) need > 168 registers per thread and
 * at most 5 of their 40 KB environments fit an SM anyway: bound that instantiation at 6 warps so that nothing spills
 * (spilled, it ran 1.9x slower: local memory competes for the ~28 KB of L1 left beside the 227 KB of shared memory). */
#define RR_MAX_WPB_WIDE 6
template <int NS> struct RRMaxWpb { static constexpr int value = NS <= 3 ? RR_MAX_WPB : (RR_MAX_WPB < RR_MAX_WPB_WIDE ? RR_MAX_WPB : RR_MAX_WPB_WIDE); };

template <int NS, bool DBG>
__global__ void __launch_bounds__(32 * RRMaxWpb<NS>::value, 1) rr_step_kernel(const __grid_constant__ RRModelDev m,
                                                                     const __grid_constant__ RRStepArgs a) {
  extern __shared__ float4 rr_smem4[];
  int32_t *ti = reinterpret_cast<int32_t *>(rr_smem4);
  float *tf = reinterpret_cast<float *>(ti + m.ni);
  float *envs = tf + m.nf;
  {
    int4 *dst = reinterpret_cast<int4 *>(ti);
    const int4 *src = reinterpret_cast<const int4 *>(m.ibuf);
    for (int i = threadIdx.x; i < m.ni / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    float4 *dstf = reinterpret_cast<float4 *>(tf);
    const float4 *srcf = reinterpret_cast<const float4 *>(m.fbuf);
    for (int i = threadIdx.x; i < m.nf / 4; i += blockDim.x) dstf[i] = __ldg(srcf + i);
  }
  __syncthreads();
  const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *sm = envs + (size_t)warp * m.sm.total;
  /* every warp of every CTA runs the same number of passes (the kernel rendezvous CTA-wide for instruction-cache
   * locality); passes beyond the batch are padding: they recompute the last environment and store nothing */
  const int stride = gridDim.x * wpb, trips = (a.B + stride - 1) / stride;
  /* without an explicit order, CTA b owns the contiguous range [b B / grid, (b + 1) B / grid): every CTA gets the same
   * number of environments (+-1), so the short last pass is spread over all SMs instead of leaving whole CTAs idle */
  const int e_beg = (int)(((long long)blockIdx.x * a.B) / gridDim.x), e_end = (int)(((long long)(blockIdx.x + 1) * a.B) / gridDim.x);
  /* When the CTA's last pass would be less than half full, its environments are spread evenly over the passes instead
   * (14 environments in 2 passes: 7 + 7 warps rather than 10 + 4): a pass costs about the same whatever its width, but a
   * little less when narrower (2048 envs: +2.3 %; 4096 envs, 10 + 10 + 8: no gain, left as is). */
  const int n_c = e_end - e_beg, last = n_c - (trips - 1) * wpb;
  const int chunk = (trips > 1 && 2 * last < wpb) ? (n_c + trips - 1) / trips : wpb;
  for (int it = 0; it < trips; it++) {
    const int slot = it * stride + blockIdx.x * wpb + warp;
    int env = e_beg + it * chunk + warp;
    if (warp >= chunk || env >= e_end) env = a.B; /* padding pass */
    if (a.env_order) { env = a.env_order[slot]; if (env < 0) env = a.B; } /* idle slot -> padding pass */
    if (DBG) rr_dbg::env_run<NS>(m, a, env, blockIdx.x * wpb + warp, sm, ti, tf, lane);
    else rr::env_run<NS>(m, a, env, blockIdx.x * wpb + warp, sm, ti, tf, lane);
    __syncwarp();
  }
}

/* Programmatic dependent launch for the learner's small kernels: each starts with RR_PDL_PROLOGUE (wait for everything the stream
 * produced, then let the next kernel of the stream begin its own launch / prologue), and is launched with the stream-serialisation
 * attribute, which hides the launch latency between the ~18 dependent launches of a minibatch update.  RR_TC_NO_PDL=1: plain launches. */
#define RR_PDL_PROLOGUE()                                      \
  do {                                                         \
    asm volatile("griddepcontrol.wait;" ::: "memory");         \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)
template <typename... KArgs, typename... Args>
static cudaError_t rr_launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, void *stream, Args... args) {
  static const bool pdl = getenv("RR_TC_NO_PDL") == nullptr && getenv("RR_TC_NO_PDL_SMALL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

/* ppo.losses.compute_gae: one thread per environment, reverse scan over the unroll (T = 10 in the reference). */
__global__ void rr_gae_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
                              const float *__restrict__ bootstrap, const float *__restrict__ termination,
                              const float *__restrict__ truncation, int T, int B, float discount, float lambda_,
                              float *__restrict__ vs, float *__restrict__ adv) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
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

/* ---- fused PPO loss (rr_ppo_loss.h): stage A one thread per environment column, stage B one thread per (t, env) element ---- */
#include "rr_ppo_loss.h"
#define RR_PPO_THREADS 128
__global__ void __launch_bounds__(RR_PPO_THREADS) rr_ppo_loss_a_kernel(const __grid_constant__ RRPpoLossArgs a) {
  __shared__ double sh1[RR_PPO_THREADS], sh2[RR_PPO_THREADS];
  RR_PDL_PROLOGUE();
  const int b = blockIdx.x * RR_PPO_THREADS + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (b < a.B) rr_ppo_stage_a(a, b, s1, s2);
  sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = RR_PPO_THREADS / 2; o > 0; o >>= 1) { /* fixed-order tree: deterministic */
    if (threadIdx.x < o) { sh1[threadIdx.x] += sh1[threadIdx.x + o]; sh2[threadIdx.x] += sh2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { a.adv_partial[2 * blockIdx.x] = sh1[0]; a.adv_partial[2 * blockIdx.x + 1] = sh2[0]; }
}
__global__ void __launch_bounds__(RR_PPO_THREADS) rr_ppo_loss_b_kernel(const __grid_constant__ RRPpoLossArgs a) {
  __shared__ float sh[3][RR_PPO_THREADS];
  RR_PDL_PROLOGUE();
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < a.nblkA; k++) { s1 += a.adv_partial[2 * k]; s2 += a.adv_partial[2 * k + 1]; }
  const double n = (double)a.T * (double)a.B, mean = s1 / n;
  double var = s2 / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const size_t i = (size_t)blockIdx.x * RR_PPO_THREADS + threadIdx.x;
  float pol = 0.f, val = 0.f, ent = 0.f;
  if (i < (size_t)a.T * a.B) rr_ppo_stage_b(a, i, (float)mean, (float)sqrt(var), pol, val, ent);
  sh[0][threadIdx.x] = pol; sh[1][threadIdx.x] = val; sh[2][threadIdx.x] = ent;
  __syncthreads();
  for (int o = RR_PPO_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int q = 0; q < 3; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 3) a.loss_partial[3 * blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}

/* FP32 FMA peak: 8 independent chains per thread (latency 4 cycles x 2 issue ports needs >= 8 in flight per thread pair) */
__global__ void __launch_bounds__(256) rr_fma_peak_kernel(float *out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.9999f, c = 1e-4f;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

static thread_local char g_cuda_err[256];
static const char *rrb_error() { return g_cuda_err; }
static int rrb_check(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return 1;
}
static int rrb_set_device(int device) { return rrb_check(cudaSetDevice(device), "cudaSetDevice"); }
static int rrb_malloc(void **p, size_t bytes) { return rrb_check(cudaMalloc(p, bytes ? bytes : 4), "cudaMalloc"); }
static void rrb_free(void *p) { cudaFree(p); }
static int rrb_h2d(void *dst, const void *src, size_t bytes, void *stream) {
  return rrb_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpyAsync H2D");
}
static int rrb_d2h(void *dst, const void *src, size_t bytes, void *stream) {
  return rrb_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync D2H");
}
static int rrb_num_slots() {
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  return (n_sm > 0 ? n_sm : 160) * RR_MAX_WPB;
}
static void *rrb_host_devptr(void *host) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}
static int rrb_sync(void *stream) { return rrb_check(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

static int rrb_fp32_peak(double *tflops, void *stream) {
  int dev = 0, n_sm = 0;
  if (rrb_check(cudaGetDevice(&dev), "cudaGetDevice") ||
      rrb_check(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute"))
    return 1;
  const int ctas = n_sm * 8, threads = 256, iters = 4096;
  float *out = nullptr;
  if (rrb_check(cudaMalloc(&out, (size_t)ctas * threads * sizeof(float)), "cudaMalloc")) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaStream_t st = (cudaStream_t)stream;
  rr_fma_peak_kernel<<<ctas, threads, 0, st>>>(out, 64);
  double best = 0.0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0, st);
    rr_fma_peak_kernel<<<ctas, threads, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * 16 * (double)iters * ctas * threads;
    if (ms > 0.f) best = best > flops / (ms * 1e-3) / 1e12 ? best : flops / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return rrb_check(cudaGetLastError(), "rr_fma_peak_kernel");
}

static int rrb_max_wpb(const RRModelDev &m) { return m.nv <= 96 ? RRMaxWpb<3>::value : RRMaxWpb<5>::value; }

static int rrb_geometry(const RRModelDev &m, int B, int *ctas, int *wpb_out) {
  const size_t tables = ((size_t)m.ni + m.nf) * 4, per_env = (size_t)m.sm.total * sizeof(float);
  if (tables + per_env > RR_SMEM_MAX) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "model needs %zu B of shared memory per environment (+%zu B tables) > %d", per_env,
             tables, RR_SMEM_MAX);
    return 1;
  }
  int wpb = (int)((RR_SMEM_MAX - tables) / per_env);
  if (wpb > rrb_max_wpb(m)) wpb = rrb_max_wpb(m);
  if (const char *ov = getenv("RR_WPB")) { /* developer knob: fewer environments per CTA (occupancy experiments) */
    int v = atoi(ov);
    if (v >= 1 && v < wpb) wpb = v;
  }
  int dev = 0, n_sm = 0;
  if (rrb_check(cudaGetDevice(&dev), "cudaGetDevice") ||
      rrb_check(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute"))
    return 1;
  int grid = (B + wpb - 1) / wpb;
  if (grid > n_sm) grid = n_sm;
  *ctas = grid;
  *wpb_out = wpb;
  return 0;
}

template <int NS, bool DBG>
static int rrb_launch_ns(const RRModelDev &m, const RRStepArgs &a, void *stream) {
  int grid = 1, wpb = 1;
  if (rrb_geometry(m, a.B, &grid, &wpb)) return 1;
  const size_t smem = ((size_t)m.ni + m.nf) * 4 + (size_t)m.sm.total * sizeof(float) * wpb;
  static std::atomic<bool> configured[64]; /* per device; the attribute call is idempotent, so a race only repeats it */
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!configured[dev].load(std::memory_order_acquire)) {
    if (rrb_check(cudaFuncSetAttribute(rr_step_kernel<NS, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, RR_SMEM_MAX),
                  "cudaFuncSetAttribute(smem)"))
      return 1;
    configured[dev].store(true, std::memory_order_release);
  }
  rr_step_kernel<NS, DBG><<<grid, 32 * wpb, smem, (cudaStream_t)stream>>>(m, a);
  return rrb_check(cudaGetLastError(), "rr_step_kernel launch");
}

static int rrb_launch_step(const RRModelDev &m, const RRStepArgs &a, void *stream) {
  const bool dbg = a.dbg.buf != nullptr || a.prof != nullptr;
  if (m.nv <= 96) return dbg ? rrb_launch_ns<3, true>(m, a, stream) : rrb_launch_ns<3, false>(m, a, stream);
  return dbg ? rrb_launch_ns<5, true>(m, a, stream) : rrb_launch_ns<5, false>(m, a, stream);
}

static int rrb_launch_gae(const float *rewards, const float *values, const float *bootstrap, const float *termination,
                          const float *truncation, int T, int B, float discount, float lambda_, float *vs, float *adv,
                          void *stream) {
  rr_gae_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, values, bootstrap, termination, truncation, T, B,
                                                                  discount, lambda_, vs, adv);
  return rrb_check(cudaGetLastError(), "rr_gae_kernel launch");
}

/* Stage B with one WARP per (t, env) element and one lane per action (A <= 32): the thread-per-element kernel above walks the
 * 30 actions twice with log / exp / tanh in every step on only 40 CTAs (33 us for the README minibatch); the same arithmetic
 * spread over the lanes, the two sums over the actions by shuffles. */
#define RR_PPO_WARPS 8
__global__ void __launch_bounds__(32 * RR_PPO_WARPS) rr_ppo_loss_b_warp_kernel(const __grid_constant__ RRPpoLossArgs a) {
  __shared__ float sh[3][RR_PPO_WARPS];
  RR_PDL_PROLOGUE();
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < a.nblkA; k++) { s1 += a.adv_partial[2 * k]; s2 += a.adv_partial[2 * k + 1]; }
  const double n = (double)a.T * (double)a.B, mean_d = s1 / n;
  double var = s2 / n - mean_d * mean_d;
  if (var < 0.0) var = 0.0;
  const float mean = (float)mean_d, std_ = (float)sqrt(var);
  const int warp = threadIdx.x >> 5, k = threadIdx.x & 31, A = a.A;
  const size_t i = (size_t)blockIdx.x * RR_PPO_WARPS + warp;
  float pol = 0.f, val = 0.f, ent = 0.f;
  if (i < (size_t)a.T * a.B) {
    const float invN = 1.f / ((float)a.T * (float)a.B);
    const float adv = a.normalize_advantage ? (a.adv[i] - mean) / (std_ + 1e-8f) : a.adv[i];
    const float *lg = a.logits + i * 2 * A, *raw = a.raw_action + i * A, *nz = a.noise + i * A;
    const bool on = k < A;
    const float loc = on ? lg[k] : 0.f, pre = on ? lg[A + k] : 0.f, rk = on ? raw[k] : 0.f, nk = on ? nz[k] : 0.f;
    const float scale = rr_softplus(pre) + 1e-3f, is = 1.f / scale, z = (rk - loc) * is, lsc = logf(scale);
    float lp = on ? -0.5f * z * z - lsc - RR_PPO_HALF_LOG_2PI - rr_log_det_tanh(rk) : 0.f;
    const float raw_e = loc + scale * nk, th = tanhf(raw_e);
    float e = on ? 0.5f + RR_PPO_HALF_LOG_2PI + lsc + rr_log_det_tanh(raw_e) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lp += __shfl_xor_sync(0xffffffffu, lp, o);
      e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    const float rho = expf(lp - a.old_log_prob[i]);
    const float lo = 1.f - a.clip_eps, hi = 1.f + a.clip_eps;
    const float t1 = rho * adv, t2 = fminf(fmaxf(rho, lo), hi) * adv;
    const float g = (t1 <= t2) ? -adv * rho * invN : 0.f;
    const float d = a.vs[i] - a.baseline[i];
    const float ce = -a.entropy_cost * invN;
    if (on) {
      const float sig = pre > 20.f ? 1.f : 1.f / (1.f + expf(-pre));
      float *gl = a.grad_logits + i * 2 * A;
      gl[k] = g * z * is + ce * (-2.f * th);
      gl[A + k] = (g * (z * z - 1.f) * is + ce * (is - 2.f * th * nk)) * sig;
    }
    if (k == 0) {
      a.lp[i] = lp;
      a.grad_baseline[i] = -0.5f * d * invN;
    }
    pol = -fminf(t1, t2); val = 0.25f * d * d; ent = e;
  }
  if (k == 0) { sh[0][warp] = pol; sh[1][warp] = val; sh[2][warp] = ent; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float acc = 0.f;
    for (int w = 0; w < RR_PPO_WARPS; w++) acc += sh[threadIdx.x][w];
    a.loss_partial[3 * blockIdx.x + threadIdx.x] = acc;
  }
}

static int rrb_ppo_blocks(int n) { return (n + RR_PPO_THREADS - 1) / RR_PPO_THREADS; }
/* stage-B partial-sum rows the caller provides: the warp-per-element kernel's block count (>= the thread-per-element one's) */
static int rrb_ppo_blocks_b(int n) { return (n + RR_PPO_WARPS - 1) / RR_PPO_WARPS; }
static int rrb_launch_ppo_loss(const RRPpoLossArgs &a, void *stream) {
  if (rrb_check(rr_launch_pdl(rr_ppo_loss_a_kernel, rrb_ppo_blocks(a.B), RR_PPO_THREADS, 0, stream, a), "rr_ppo_loss_a launch")) return 1;
  if (a.A <= 32) {
    if (rrb_check(rr_launch_pdl(rr_ppo_loss_b_warp_kernel, rrb_ppo_blocks_b(a.T * a.B), 32 * RR_PPO_WARPS, 0, stream, a), "rr_ppo_loss_b launch"))
      return 1;
  } else { /* wide action spaces: thread per element; the partial-sum rows it does not write are zeroed */
    const int used = rrb_ppo_blocks(a.T * a.B), rows = rrb_ppo_blocks_b(a.T * a.B);
    if (rows > used) cudaMemsetAsync(a.loss_partial + 3 * (size_t)used, 0, sizeof(float) * 3 * (size_t)(rows - used), (cudaStream_t)stream);
    rr_ppo_loss_b_kernel<<<used, RR_PPO_THREADS, 0, (cudaStream_t)stream>>>(a);
  }
  return rrb_check(cudaGetLastError(), "rr_ppo_loss launch");
}

/* ---- Adam on the flat parameter buffer, minibatch gather (rr_learner_misc.h) ---- */
#define RR_MISC_HD __host__ __device__ static inline
#include "rr_learner_misc.h"
__global__ void rr_adam_kernel(float *__restrict__ p, float *__restrict__ g, const float *__restrict__ partials, int nsplit,
                               float *__restrict__ m, float *__restrict__ v, const float *__restrict__ step, long long n, float lr,
                               float b1, float b2, float eps) {
  RR_PDL_PROLOGUE();
  const float t = step[0] + 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi;
    if (partials) { /* the split weight-gradient launch's partial sums, added in a fixed order */
      gi = partials[i];
      for (int s = 1; s < nsplit; s++) gi += partials[(size_t)s * n + i];
      g[i] = gi;
    } else gi = g[i];
    float pi = p[i], mi = m[i], vi = v[i];
    rr_adam_element(pi, gi, mi, vi, t, lr, b1, b2, eps);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}
__global__ void rr_adam_step_inc_kernel(float *step) {
  RR_PDL_PROLOGUE();
  step[0] += 1.f;
}
static int rrb_adam_step(float *p, float *g, const float *partials, int nsplit, float *m, float *v, float *step, long long n, float lr,
                         float b1, float b2, float eps, void *stream) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (rrb_check(rr_launch_pdl(rr_adam_kernel, (int)blocks, 256, 0, stream, p, g, partials, nsplit, m, v, (const float *)step, n, lr, b1, b2, eps), "rr_adam_kernel launch"))
    return 1;
  return rrb_check(rr_launch_pdl(rr_adam_step_inc_kernel, 1, 1, 0, stream, step), "rr_adam_step_inc_kernel launch");
}
/* one block per gathered row: dst[t, j, :] = src[t, idx[j], :] */
__global__ void rr_gather_kernel(const __grid_constant__ RRGatherArgs a) {
  RR_PDL_PROLOGUE();
  int it = 0;
  for (int i = 1; i < a.count; i++)
    if ((int)blockIdx.x >= a.block_start[i]) it = i;
  const rr_gather_item g = a.item[it];
  const int r = blockIdx.x - a.block_start[it], t = r / a.rows, j = r % a.rows;
  const long long src_row = a.idx[j];
  const float *src = g.src + ((size_t)t * g.src_rows + src_row) * g.inner;
  float *dst = g.dst + ((size_t)t * a.rows + j) * g.dst_pitch;
  if (((g.inner | g.dst_pitch) & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.src) | reinterpret_cast<uintptr_t>(g.dst)) & 15) == 0) {
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = threadIdx.x; i < g.inner / 4; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (int i = threadIdx.x; i < g.inner; i += blockDim.x) dst[i] = src[i];
  }
}
static int rrb_gather_rows(const RRGatherArgs &a, int blocks, void *stream) {
  return rrb_check(rr_launch_pdl(rr_gather_kernel, blocks, 128, 0, stream, a), "rr_gather_kernel launch");
}

/* ---- policy inference of the rollout: normalise -> MLP (32-wide swish layers) -> tanh-normal sample + log-prob, one kernel ----
 * One warp per environment row, lane j = neuron j.  All weights are staged once per (persistent) CTA in shared memory with odd row
 * pitches, so that "lane j reads column k of its own row" is conflict-free; the input / activations are broadcast by shuffles. */
#define RR_POLICY_WARPS 8
__device__ __forceinline__ int rr_policy_pitch0(int obs_dim) { return (((obs_dim + 31) / 32) * 32) | 1; }
__global__ void __launch_bounds__(32 * RR_POLICY_WARPS, 1) rr_policy_act_kernel(const __grid_constant__ rr_policy_args a) {
  extern __shared__ float ps[];
  RR_PDL_PROLOGUE();
  const int H = RR_POLICY_HIDDEN, P0 = rr_policy_pitch0(a.obs_dim), PH = H + 1, nh = a.nlayers - 2, A = a.A;
  float *w0 = ps, *wh = w0 + H * P0, *wo = wh + nh * H * PH, *ms = wo + 2 * A * PH, *sd = ms + a.obs_dim;
  for (int i = threadIdx.x; i < H * P0; i += blockDim.x) {
    const int j = i / P0, k = i - j * P0;
    w0[i] = k < a.obs_dim ? a.w[0][(size_t)j * a.in0 + k] : 0.f;
  }
  for (int l = 0; l < nh; l++)
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) wh[l * H * PH + (i / H) * PH + (i % H)] = a.w[1 + l][i];
  for (int i = threadIdx.x; i < 2 * A * H; i += blockDim.x) wo[(i / H) * PH + (i % H)] = a.w[a.nlayers - 1][i];
  for (int i = threadIdx.x; i < a.obs_dim; i += blockDim.x) {
    ms[i] = a.mean ? a.mean[i] : 0.f;
    sd[i] = a.mean ? a.std[i] : 1.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * RR_POLICY_WARPS + warp; r < a.B; r += gridDim.x * RR_POLICY_WARPS) {
    /* layer 0: obs_dim -> 32 */
    const float *x = a.obs + (size_t)r * a.obs_dim, *wrow = w0 + lane * P0;
    float acc[4] = {__ldg(a.b[0] + lane), 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < a.obs_dim; k0 += 32) {
      const int k = k0 + lane;
      float xv = 0.f;
      if (k < a.obs_dim) xv = (x[k] - ms[k]) / sd[k];
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] += __shfl_sync(0xffffffffu, xv, kk + u) * wrow[k0 + kk + u];
      }
    }
    float h = rr_policy_silu((acc[0] + acc[1]) + (acc[2] + acc[3]));
    /* hidden layers 32 -> 32 */
    for (int l = 0; l < nh; l++) {
      const float *wr = wh + l * H * PH + lane * PH;
      float c[4] = {__ldg(a.b[1 + l] + lane), 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < 32; kk += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) c[u] += __shfl_sync(0xffffffffu, h, kk + u) * wr[kk + u];
      }
      h = rr_policy_silu((c[0] + c[1]) + (c[2] + c[3]));
    }
    /* head 32 -> 2 A: lane k < A forms loc_k and the pre-softplus scale of action k */
    const bool on = lane < A;
    const float *wa = wo + (on ? lane : 0) * PH, *wb = wo + (on ? A + lane : 0) * PH;
    const float *bo = a.b[a.nlayers - 1];
    float loc = on ? __ldg(bo + lane) : 0.f, pre = on ? __ldg(bo + A + lane) : 0.f;
#pragma unroll
    for (int kk = 0; kk < 32; kk++) {
      const float hv = __shfl_sync(0xffffffffu, h, kk);
      loc += hv * wa[kk];
      pre += hv * wb[kk];
    }
    float lp = 0.f;
    if (on) {
      float act, raw;
      lp = rr_policy_sample(loc, pre, a.eps ? a.eps + (size_t)r * A + lane : nullptr, act, raw);
      a.action[(size_t)r * A + lane] = act;
      a.raw_action[(size_t)r * A + lane] = raw;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
    if (lane == 0) a.log_prob[r] = lp;
  }
}
static int rrb_policy_act(const rr_policy_args &a, void *stream) {
  const int H = RR_POLICY_HIDDEN, P0 = (((a.obs_dim + 31) / 32) * 32) | 1;
  const size_t smem = sizeof(float) * ((size_t)H * P0 + (size_t)(a.nlayers - 2) * H * (H + 1) + (size_t)2 * a.A * (H + 1) + 2 * (size_t)a.obs_dim);
  if (smem > RR_SMEM_MAX) { snprintf(g_cuda_err, sizeof(g_cuda_err), "rr_policy_act: %zu bytes of weights do not fit shared memory", smem); return 1; }
  static std::atomic<bool> configured[64];
  int dev = 0;
  if (rrb_check(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
  dev &= 63;
  if (!configured[dev].load(std::memory_order_acquire)) {
    if (rrb_check(cudaFuncSetAttribute(rr_policy_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RR_SMEM_MAX), "rr_policy_act attribute"))
      return 1;
    configured[dev].store(true, std::memory_order_release);
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = (a.B + RR_POLICY_WARPS - 1) / RR_POLICY_WARPS;
  if (grid > sms) grid = sms;
  return rrb_check(rr_launch_pdl(rr_policy_act_kernel, grid, 32 * RR_POLICY_WARPS, smem, stream, a), "rr_policy_act_kernel launch");
}

/* grouped TF32 GEMM of the learner on the tensor cores (tcgen05) */
#define RR_TC_HD __host__ __device__ static inline
#include "rr_tc_gemm.h"
#define RR_TC_BN_MAX 128
static int rrb_tc_smem_max() { return RR_TC_STAGES * (RR_TC_BM + RR_TC_BN_MAX) * RR_TC_BK * 4; }
/* Tensor maps for the operands TMA can fetch: 16-byte aligned base and pitch.  cuTensorMapEncodeTiled comes from the driver
 * through the runtime's entry-point query (the library does not link libcuda); without it every operand takes the cp.async path. */
typedef CUresult (*rr_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static rr_encode_tiled_fn rrb_encode_fn() {
  static std::atomic<void *> cached{nullptr};
  static std::atomic<bool> tried{false};
  if (!tried.load(std::memory_order_acquire)) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    cudaGetLastError();
    cached.store(fn, std::memory_order_release);
    tried.store(true, std::memory_order_release);
  }
  return (rr_encode_tiled_fn)cached.load(std::memory_order_acquire);
}
/* operand stored row-major with pitch ld: K-major = [rows, k] (box 32 k x tile rows, 128-byte swizzle); MN-major = [k, rows]
 * (box 32 rows x 32 k, 128-byte swizzle with 32-byte atoms) */
static bool rrb_tc_encode_operand(uint64_t *out, const float *base, int rows, int k, int ld, int mn, int tile_rows) {
  rr_encode_tiled_fn fn = rrb_encode_fn();
  if (!fn || (reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4) != 0 || tile_rows > 256) return false;
  alignas(64) CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)(mn ? rows : k), (cuuint64_t)(mn ? k : rows)};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)(mn ? 32 : tile_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  memcpy(out, &map, sizeof(map));
  return true;
}
static void rrb_tc_encode(RRTcRecord &rec) {
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  rr_tc_problem &p = rec.p;
  int flags = 0;
  if (getenv("RR_TC_NO_TMA") == nullptr) {
    if (rrb_tc_encode_operand(rec.tmap_a, p.a, p.m, p.k, p.lda, p.a_mn, RR_TC_BM)) flags |= 1;
    if (rrb_tc_encode_operand(rec.tmap_b, p.b, p.n, p.k, p.ldb, p.b_mn, p.bn)) flags |= 2;
  }
  p.reserved[2] = flags;
}
static int rrb_tc_launch(const RRTcRecord *dev_recs, int count, int total_tiles, int smem_bytes, void *stream) {
  static std::atomic<bool> configured[64];
  int dev = 0;
  if (rrb_check(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
  dev &= 63;
  if (!configured[dev].load(std::memory_order_acquire)) {
    if (rrb_check(cudaFuncSetAttribute(rr_tc::gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rrb_tc_smem_max()),
                  "rr_tc gemm_kernel attribute"))
      return 1;
    configured[dev].store(true, std::memory_order_release);
  }
  /* Programmatic dependent launch: the kernel's prologue (problem lookup, TMEM allocation, mbarrier initialisation) may start
   * while the previous kernel of the stream drains; it executes griddepcontrol.wait before it touches anything the stream
   * produced.  RR_TC_NO_PDL=1 launches plainly. */
  static const bool pdl = getenv("RR_TC_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)total_tiles);
  cfg.blockDim = dim3(RR_TC_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return rrb_check(cudaLaunchKernelEx(&cfg, rr_tc::gemm_kernel, dev_recs, count), "rr_tc gemm_kernel launch");
}

#include "rr_api_impl.inl"
