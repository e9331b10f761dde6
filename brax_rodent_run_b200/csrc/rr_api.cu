/* rr_api.cu -- CUDA (sm_100a) backend of the C ABI in include/rr_b200.h.
 *
 * Build (see brax_rodent_run_b200/build.py):
 *   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC -o librr_b200.so rr_api.cu
 *
 * Launch geometry: one warp per environment, RR_WPB warps per CTA; each warp owns `sm.total` floats of dynamic
 * shared memory holding the whole per-environment working set of a physics substep (about 33 KB for
 * rodent_0.xml), so HBM is touched only to load the 1 KB state + action and to store state + observation.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <math_constants.h>

#include "rr_kernels.inl"

#ifndef RR_WPB
#define RR_WPB 1
#endif

template <int NS>
__global__ void __launch_bounds__(32 * RR_WPB) rr_step_kernel(const __grid_constant__ RRModelDev m,
                                                              const __grid_constant__ RRStepArgs a) {
  extern __shared__ float4 rr_smem4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * RR_WPB + warp;
  if (env >= a.B) return;
  float *sm = reinterpret_cast<float *>(rr_smem4) + (size_t)warp * m.sm.total;
  rr::env_run<NS>(m, a, env, sm, lane);
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
static int rrb_sync(void *stream) { return rrb_check(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

template <int NS>
static int rrb_launch_ns(const RRModelDev &m, const RRStepArgs &a, void *stream) {
  size_t smem = (size_t)m.sm.total * sizeof(float) * RR_WPB;
  static thread_local size_t configured = 0;
  if (smem > configured) {
    if (rrb_check(cudaFuncSetAttribute(rr_step_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "cudaFuncSetAttribute(smem)"))
      return 1;
    cudaFuncSetAttribute(rr_step_kernel<NS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = smem;
  }
  int grid = (a.B + RR_WPB - 1) / RR_WPB;
  rr_step_kernel<NS><<<grid, 32 * RR_WPB, smem, (cudaStream_t)stream>>>(m, a);
  return rrb_check(cudaGetLastError(), "rr_step_kernel launch");
}

static int rrb_launch_step(const RRModelDev &m, const RRStepArgs &a, void *stream) {
  if (m.nv <= 96) return rrb_launch_ns<3>(m, a, stream);
  return rrb_launch_ns<5>(m, a, stream);
}

static int rrb_launch_gae(const float *rewards, const float *values, const float *bootstrap, const float *termination,
                          const float *truncation, int T, int B, float discount, float lambda_, float *vs, float *adv,
                          void *stream) {
  rr_gae_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, values, bootstrap, termination, truncation, T, B,
                                                                  discount, lambda_, vs, adv);
  return rrb_check(cudaGetLastError(), "rr_gae_kernel launch");
}

#include "rr_api_impl.inl"
