// Latency microbenchmarks (one warp): dependent chains of the primitives the step kernel's serial paths are made of.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float *out, long long *t, int n) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-3f * (i % 7);
  __syncthreads();
  int lane = threadIdx.x & 31;
  float x = 1.f + lane, y = 0.5f;
  long long t0, t1;
  // (a) shfl -> fma chain
  t0 = clock64();
  for (int i = 0; i < n; i++) { float v = __shfl_sync(0xffffffffu, x, i & 31); x = fmaf(-1e-3f, v, x); }
  t1 = clock64(); if (threadIdx.x == 0) t[0] = t1 - t0;
  // (b) dependent LDS chain (address from previous value)
  int a = lane;
  t0 = clock64();
  for (int i = 0; i < n; i++) { float v = sm[a & 4095]; a = a + 1 + (int)(v * 0.f); y += v; }
  t1 = clock64(); if (threadIdx.x == 0) t[1] = t1 - t0;
  // (b2) truly dependent lds: address depends on loaded value
  int b = lane; 
  int *smi = (int*)sm;
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) smi[i] = (i * 17 + 5) & 4095;
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < n; i++) { b = smi[b]; }
  t1 = clock64(); if (threadIdx.x == 0) t[2] = t1 - t0;
  // (c) rcp chain
  float r = 1.5f + lane;
  t0 = clock64();
  for (int i = 0; i < n; i++) { float q; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(r)); r = q + 1.0f; }
  t1 = clock64(); if (threadIdx.x == 0) t[3] = t1 - t0;
  // (d) fma chain
  float f = x;
  t0 = clock64();
  for (int i = 0; i < n; i++) { f = fmaf(f, 0.999f, 1e-3f); }
  t1 = clock64(); if (threadIdx.x == 0) t[4] = t1 - t0;
  // (e) shfl_xor reduction round chain (5 rounds) per iteration
  float g = f;
  t0 = clock64();
  for (int i = 0; i < n; i++) { for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o); g *= 1e-3f; }
  t1 = clock64(); if (threadIdx.x == 0) t[5] = t1 - t0;
  // (f) sts -> syncwarp -> lds round trip
  float h = g;
  t0 = clock64();
  for (int i = 0; i < n; i++) { sm[lane] = h; __syncwarp(); h = sm[(lane + 1) & 31] + 1e-3f; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) t[6] = t1 - t0;
  // (g) constant-bank indexed load chain is not expressible simply; ballot+popc chain
  unsigned m = 0x55555555u;
  t0 = clock64();
  for (int i = 0; i < n; i++) { unsigned bb = __ballot_sync(0xffffffffu, (m >> lane) & 1); m = bb * 3u + __popc(bb); }
  t1 = clock64(); if (threadIdx.x == 0) t[7] = t1 - t0;
  out[threadIdx.x] = x + y + b + r + f + g + h + m;
}
int main() {
  float *out; long long *t; cudaMalloc(&out, 4096); cudaMalloc(&t, 64);
  const char *names[] = {"shfl.idx->ffma", "lds (indep addr)+fadd", "lds->lds dependent", "mufu.rcp->fadd", "ffma", "5x(shfl.xor+fadd)+fmul", "sts->syncwarp->lds->fadd->syncwarp", "ballot->imad/popc"};
  for (int warps = 1; warps <= 8; warps *= 8) {
    int n = 2000;
    k<<<1, 32 * warps>>>(out, t, n); cudaDeviceSynchronize();
    k<<<1, 32 * warps>>>(out, t, n); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
    printf("warps in CTA: %d\n", warps);
    for (int i = 0; i < 8; i++) printf("  %-40s %.1f cycles/iter\n", names[i], (double)h[i] / n);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
