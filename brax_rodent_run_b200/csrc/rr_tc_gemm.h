/* rr_tc_gemm.h -- grouped TF32 GEMM on the Blackwell tensor cores (tcgen05.mma kind::tf32, accumulator in tensor memory) with
 * the learner's epilogues fused (bias, SiLU, SiLU-derivative scaling, bias-gradient column).  See rr_tc_problem in
 * include/rr_b200.h for the semantics; the host-side reference of the same semantics is rr_tc_reference() below (used by the
 * emulator backend under tests/emu so that the Python learner's problem lists can be checked on the CPU).
 *
 * One CTA (256 threads) owns one 128 x bn output tile:
 *   - operand k-blocks (32 wide) are staged in the UMMA canonical SWIZZLED shared-memory layouts, which are also what TMA writes:
 *         K-major operand (layout type 2, 128-byte swizzle): row r of the tile is one 128-byte line at (r / 8) * 1024 +
 *             (r % 8) * 128, its 16-byte chunk c stored at chunk c ^ (r % 8)                          [SBO 1024; k-step +32 B]
 *         MN-major operand (layout type 1, 128-byte swizzle with 32-byte atomicity -- the only layout the hardware takes for
 *             32-bit MN-major operands): k-line k of a 32-row block is one 128-byte line at block * 4096 + k * 128, its
 *             32-byte unit u stored at unit u ^ (k % 4)                                        [LBO 4096, SBO 512; k-step +1024 B]
 *     so row-major X, W (K-major for X W') and the same arrays read "transposed" (MN-major, for dY W and dY' X) need no copy;
 *   - RR_TC_STAGES-deep ring, warp-specialised: producer warp w owns stage w.  Operands that a tensor map can describe
 *     (16-byte aligned base and pitch) arrive by TMA (cp.async.bulk.tensor.2d, complete_tx on the stage's "full" mbarrier;
 *     out-of-range rows / k are zero-filled by the TMA unit); the others -- unaligned operands such as the value head's, and the
 *     tile that carries the virtual row of ones (bias gradient) -- by cp.async (16 bytes, or 4 with explicit zero fill) followed
 *     by cp.async.wait_group 0, fence.proxy.async and an arrive on "full".  One thread of warp 4 waits for "full", issues the
 *     four tcgen05.mma (K = 8 each) of the stage and tcgen05.commit's to the stage's "empty" mbarrier, which the producer waits
 *     on before it overwrites the stage;
 *   - epilogue: each warp reads its 32 TMEM lanes (= rows) with tcgen05.ld 32x32b.x16, applies the epilogue and stores rows.
 */
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/rr_b200.h"

#define RR_TC_BM 128
#define RR_TC_BK 32
#define RR_TC_STAGES 4
#define RR_TC_THREADS 256

/* what rr_tc_plan hands to the device per problem: the problem, then the two operands' tensor maps (CUtensorMap, 128 bytes
 * each; valid when the flag bit in p.reserved[2] is set: bit 0 = A, bit 1 = B) */
struct alignas(128) RRTcRecord {
  rr_tc_problem p;
  uint64_t tmap_a[16], tmap_b[16];
};
static_assert(sizeof(rr_tc_problem) == 128 && sizeof(RRTcRecord) == 384, "rr_tc record layout");

RR_TC_HD float rr_tc_sigmoid(float z) { return 1.f / (1.f + expf(-z)); }
RR_TC_HD float rr_tc_epilogue(const rr_tc_problem &p, int row, int col, float acc) {
  float v = acc + (p.bias ? p.bias[col] : 0.f);
  if (p.epi == 1) {
    if (p.aux_out) p.aux_out[(size_t)row * p.ldaux + col] = v;
    v = v * rr_tc_sigmoid(v);
  } else if (p.epi == 2) {
    const float z = p.aux_in[(size_t)row * p.ldaux + col], s = rr_tc_sigmoid(z);
    v *= s * (1.f + z * (1.f - s));
  }
  return v;
}

/* plain loops with the same semantics (fp32 products): emulator backend and documentation of the contract */
static inline void rr_tc_reference(const rr_tc_problem &p) {
  for (int r = 0; r < p.m; r++) {
    const int n_d = p.n - (p.b_ones == 2 ? 1 : 0); /* columns of D; column n_d (virtual or B's real last row) -> ones_out */
    for (int c = 0; c < n_d + (p.b_ones ? 1 : 0); c++) {
      float acc = 0.f;
      for (int kk = 0; kk < p.k; kk++) {
        const float av = p.a_mn ? p.a[(size_t)kk * p.lda + r] : p.a[(size_t)r * p.lda + kk];
        const float bv = (c == p.n && p.b_ones == 1) ? 1.f : (p.b_mn ? p.b[(size_t)kk * p.ldb + c] : p.b[(size_t)c * p.ldb + kk]);
        acc += av * bv;
      }
      if (c == n_d && p.b_ones) p.ones_out[r] = acc;
      else p.d[(size_t)r * p.ldd + c] = rr_tc_epilogue(p, r, c, acc);
    }
  }
}

#ifdef __CUDACC__
namespace rr_tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
/* bounded: a broken pipeline traps instead of hanging the GPU */
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0;; spin++) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t dst, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"(v) : "memory");
}
/* round to nearest TF32: outputs that are only ever read as MMA operands again (activations, their gradients) are stored
 * pre-rounded, so that the tensor core's truncation of the operand is exact and the rounding error stays unbiased */
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ void st_shared_zero16(uint32_t dst) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "f"(0.f) : "memory");
}

/* shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets in 16-byte units,
 * version 1 (Blackwell), layout type in bits 61-63 (0 no swizzle, 1 = 128-byte swizzle with 32-byte atomicity) */
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout_type << 61);
}

struct Operand {
  const float *base;
  int rows, ld, mn, row0, tile_rows, fast, ones_row; /* ones_row: global row index that reads as 1 (or -1) */
};

/* Chunk c (one 16-byte line) of an operand tile at k-block 0: shared-memory offset, global element offset, how many of its 4
 * elements are inside the matrix when the k index is (rows4; for K-major operands: 4 if the row is), the k index and the
 * position of the virtual "one" among the 4 elements (-1: none). */
struct ChunkGeom { uint32_t dst; size_t src; int rows4, k, one; };
__device__ __forceinline__ ChunkGeom chunk_geom(const Operand &o, int c) {
  ChunkGeom g;
  if (!o.mn) {
    const int l = c & 7, kc = (c >> 3) & 7, cr = c >> 6;
    const int row = o.row0 + cr * 8 + l;
    g.k = kc * 4;
    g.dst = cr * 1024 + l * 128 + ((kc ^ l) << 4);
    g.src = (size_t)row * o.ld + g.k;
    g.rows4 = row < o.rows ? 4 : 0;
    g.one = -1;
  } else {
    /* a warp's 32 consecutive chunks = one 512-byte swizzle atom: 4 k-lines of 128 contiguous bytes (32 rows) each */
    const int j = c & 7, kr = (c >> 3) & 3, rest = c >> 5, kq = rest & 7, mb = rest >> 3;
    const int r0 = o.row0 + mb * 32 + j * 4;
    g.k = kq * 4 + kr;
    g.dst = mb * 4096 + kq * 512 + kr * 128 + ((((j >> 1) ^ kr) << 5) | ((j & 1) << 4));
    g.src = (size_t)g.k * o.ld + r0;
    const int cnt = o.rows - r0;
    g.rows4 = cnt > 4 ? 4 : (cnt < 0 ? 0 : cnt);
    g.one = (o.ones_row >= r0 && o.ones_row < r0 + 4) ? o.ones_row - r0 : -1;
  }
  return g;
}
__device__ __forceinline__ int tile_chunks(const Operand &o) {
  return (o.mn ? ((o.tile_rows + 31) >> 5) * 32 : o.tile_rows) * (RR_TC_BK / 4);
}

/* one k-block of one operand tile into shared memory, any shape: ragged edges are zero-filled, unaligned operands go 4 bytes
 * at a time; the calling warp's lanes handle chunks lane, lane + 32, ... */
__device__ __forceinline__ void load_tile(const Operand &o, uint32_t sbase, int kb, int K, int lane) {
  const int nchunk = tile_chunks(o), kbase = kb * RR_TC_BK;
  const size_t kstep = (size_t)kbase * (o.mn ? o.ld : 1);
  for (int c = lane; c < nchunk; c += 32) {
    const ChunkGeom g = chunk_geom(o, c);
    const int k = kbase + g.k;
    int cnt;
    if (!o.mn) { cnt = K - k; cnt = g.rows4 ? (cnt > 4 ? 4 : cnt) : 0; }
    else cnt = k < K ? g.rows4 : 0;
    const int one = k < K ? g.one : -1;
    const uint32_t dst = sbase + g.dst;
    const float *src = o.base + g.src + kstep;
    if (cnt == 4 && o.fast) cp_async16(dst, src);
    else if (cnt <= 0 && one < 0) st_shared_zero16(dst);
    else {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        if (e < cnt) cp_async4(dst + 4 * e, src + e);
        else st_shared_f32(dst + 4 * e, e == one ? 1.f : 0.f);
      }
    }
  }
}

/* The common case -- every row of the tile inside the matrix, 16-byte aligned, the k-block entirely below K -- with the chunk
 * walk reduced to pointer increments (same chunk -> lane assignment as chunk_geom: chunk = lane + 32 i). */
__device__ __forceinline__ bool tile_interior(const Operand &o) {
  const int padded = o.mn ? ((o.tile_rows + 31) & ~31) : o.tile_rows;
  return o.fast && o.row0 + padded <= o.rows;
}
__device__ __forceinline__ void load_tile_interior(const Operand &o, uint32_t sbase, int kb, int lane) {
  const int kbase = kb * RR_TC_BK;
  if (!o.mn) {
    const int l = lane & 7, kq = lane >> 3;
    const float *src = o.base + (size_t)(o.row0 + l) * o.ld + kbase + kq * 4;
    uint32_t dst = sbase + l * 128 + ((kq ^ l) << 4);
    const size_t step = (size_t)8 * o.ld;
#pragma unroll 4
    for (int g = 0; g < o.tile_rows / 8; g++) {
      cp_async16(dst, src);
      cp_async16(dst ^ 64, src + 16); /* chunk kq + 4: (kq + 4) ^ l = (kq ^ l) ^ 4; the tile base is 1024-byte aligned */
      src += step;
      dst += 1024;
    }
  } else {
    const int j = lane & 7, kr = lane >> 3;
    const float *src0 = o.base + (size_t)(kbase + kr) * o.ld + o.row0 + j * 4;
    uint32_t dst = sbase + kr * 128 + ((((j >> 1) ^ kr) << 5) | ((j & 1) << 4));
    const size_t step = (size_t)4 * o.ld;
    for (int mb = 0; mb < (o.tile_rows + 31) / 32; mb++) {
      const float *src = src0 + mb * 32;
#pragma unroll
      for (int kq = 0; kq < 8; kq++) {
        cp_async16(dst, src);
        src += step;
        dst += 512;
      }
    }
  }
}
/* Aligned operand, k-block entirely below K, but the tile hangs over the last row (or carries the row of ones): validity is
 * per row group, so the walk is still pointer increments. */
__device__ __forceinline__ void load_tile_ragged_rows(const Operand &o, uint32_t sbase, int kb, int lane) {
  const int kbase = kb * RR_TC_BK;
  if (!o.mn) {
    const int l = lane & 7, kq = lane >> 3;
    const float *src = o.base + (size_t)(o.row0 + l) * o.ld + kbase + kq * 4;
    uint32_t dst = sbase + l * 128 + ((kq ^ l) << 4);
    const size_t step = (size_t)8 * o.ld;
    for (int g = 0; g < o.tile_rows / 8; g++) {
      if (o.row0 + g * 8 + l < o.rows) {
        cp_async16(dst, src);
        cp_async16(dst ^ 64, src + 16);
      } else {
        st_shared_zero16(dst);
        st_shared_zero16(dst ^ 64);
      }
      src += step;
      dst += 1024;
    }
  } else {
    const int j = lane & 7, kr = lane >> 3;
    const float *src0 = o.base + (size_t)(kbase + kr) * o.ld + o.row0 + j * 4;
    uint32_t dst = sbase + kr * 128 + ((((j >> 1) ^ kr) << 5) | ((j & 1) << 4));
    const size_t step = (size_t)4 * o.ld;
    for (int mb = 0; mb < (o.tile_rows + 31) / 32; mb++) {
      const float *src = src0 + mb * 32;
      const int r0 = o.row0 + mb * 32 + j * 4;
      int cnt = o.rows - r0;
      cnt = cnt > 4 ? 4 : cnt;
      const int one = (o.ones_row >= r0 && o.ones_row < r0 + 4) ? o.ones_row - r0 : -1;
      for (int kq = 0; kq < 8; kq++) {
        if (cnt == 4) cp_async16(dst, src);
        else if (cnt <= 0 && one < 0) st_shared_zero16(dst);
        else {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (e < cnt) cp_async4(dst + 4 * e, src + e);
            else st_shared_f32(dst + 4 * e, e == one ? 1.f : 0.f);
          }
        }
        src += step;
        dst += 512;
      }
    }
  }
}
__device__ __forceinline__ void load_stage(const Operand &o, bool interior, uint32_t sbase, int kb, int K, int lane) {
  if ((kb + 1) * RR_TC_BK <= K && o.fast) {
    if (interior) load_tile_interior(o, sbase, kb, lane);
    else load_tile_ragged_rows(o, sbase, kb, lane);
  } else load_tile(o, sbase, kb, K, lane);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
/* one k-block of an operand tile by TMA (issued by one thread): K-major = one box of 32 k x tile rows; MN-major = one box of
 * 32 rows x 32 k per 32-row block */
__device__ __forceinline__ void tma_load_tile(const Operand &o, const void *tmap, uint32_t sbase, int kb, uint32_t bar) {
  if (!o.mn) tma_load_2d(sbase, tmap, kb * RR_TC_BK, o.row0, bar);
  else
    for (int mb = 0; mb < (o.tile_rows + 31) / 32; mb++) tma_load_2d(sbase + mb * 4096, tmap, o.row0 + mb * 32, kb * RR_TC_BK, bar);
}

/* sigmoid with the fast exponential and reciprocal (relative error ~1e-6, far below TF32) */
__device__ __forceinline__ float sigmoid_fast(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }

__global__ void __launch_bounds__(RR_TC_THREADS, 2) gemm_kernel(const RRTcRecord *__restrict__ recs, int nprob) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ __align__(8) uint64_t full_bar[RR_TC_STAGES], empty_bar[RR_TC_STAGES], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[128]; /* this tile's slice of the bias (zeros without one) */
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  /* which problem, which tile */
  int pi = 0;
  for (int i = 1; i < nprob; i++)
    if ((int)blockIdx.x >= recs[i].p.tile_start) pi = i;
  const rr_tc_problem p = recs[pi].p;
  const int t = blockIdx.x - p.tile_start, tile_m = t / p.tiles_n, tile_n = t % p.tiles_n;
  const int BN = p.bn, K = p.k, nkb = (K + RR_TC_BK - 1) / RR_TC_BK;
  const int nst = p.reserved[3]; /* stages of this problem's ring (<= RR_TC_STAGES): what fits the launch's shared memory */
  const int n_d = p.n - (p.b_ones == 2 ? 1 : 0); /* columns of D */
  const int n_ext = n_d + (p.b_ones ? 1 : 0);   /* columns of the product: + the bias-gradient column */

  Operand oa, ob;
  oa.base = p.a; oa.rows = p.m; oa.ld = p.lda; oa.mn = p.a_mn; oa.row0 = tile_m * RR_TC_BM; oa.tile_rows = RR_TC_BM;
  oa.fast = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.a) & 15) == 0); oa.ones_row = -1;
  ob.base = p.b; ob.rows = p.n; ob.ld = p.ldb; ob.mn = p.b_mn; ob.row0 = tile_n * BN; ob.tile_rows = BN;
  ob.fast = (p.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.b) & 15) == 0); ob.ones_row = p.b_ones == 1 ? p.n : -1;

  const uint32_t smem0 = smem_u32(tc_smem);
  const uint32_t a_bytes = RR_TC_BM * RR_TC_BK * 4, b_bytes = p.b_mn ? (uint32_t)((BN + 31) >> 5) * 4096 : (uint32_t)BN * 128, stage_bytes = a_bytes + b_bytes;
  uint32_t ncols = 32;
  while ((int)ncols < BN) ncols <<= 1;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  /* TMA for an operand when the plan made a tensor map for it -- except for the B tile that carries the virtual row of ones */
  const bool tma_a = (p.reserved[2] & 1) != 0;
  const bool tma_b = (p.reserved[2] & 2) != 0 && !(p.b_ones == 1 && ob.row0 + ((BN + 31) & ~31) > p.n);
  const uint32_t tx_bytes = (tma_a ? a_bytes : 0u) + (tma_b ? b_bytes : 0u);
  const bool any_tma = tma_a || tma_b, any_cp = !tma_a || !tma_b;
  if (tid == 0) {
    for (int s = 0; s < RR_TC_STAGES; s++) {
      mbar_init(smem_u32(&full_bar[s]), (any_tma ? 1 : 0) + (any_cp ? 1 : 0)); /* expect_tx arrive + the cp.async path's arrive */
      mbar_init(smem_u32(&empty_bar[s]), 1); /* tcgen05.commit */
    }
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  /* everything above touched only the launch's own records: under programmatic dependent launch it overlaps the tail of the
   * previous kernel of the stream.  From here on the kernel reads what the stream produced. */
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid < 128) {
    const int c = tile_n * BN + tid;
    bias_s[tid] = (p.bias && tid < BN && c < n_d) ? p.bias[c] : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  /* instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A / B tf32, majors, N >> 3, M >> 4 */
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.a_mn & 1) << 15) | ((uint32_t)(p.b_mn & 1) << 16) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(RR_TC_BM >> 4) << 24);
  /* per-MMA (K = 8) descriptor geometry */
  const uint32_t a_lbo = p.a_mn ? 4096 : 16, a_sbo = p.a_mn ? 512 : 1024, a_step = p.a_mn ? 1024 : 32, a_type = p.a_mn ? 1 : 2;
  const uint32_t b_lbo = p.b_mn ? 4096 : 16, b_sbo = p.b_mn ? 512 : 1024, b_step = p.b_mn ? 1024 : 32, b_type = p.b_mn ? 1 : 2;

  /* optional per-CTA cycle counters (tools/tc_learner_timing.py --prof): reserved[0..1] = address of int64 [tiles][16] */
  long long *prof = reinterpret_cast<long long *>(((unsigned long long)(uint32_t)p.reserved[1] << 32) | (uint32_t)p.reserved[0]);
  if (prof) prof += (size_t)blockIdx.x * 16;
  const long long t_start = prof ? clock64() : 0;
  if (warp < nst) {
    /* producer warp w owns stage w: k-blocks w, w + nst, ... */
    const int s = warp;
    const bool int_a = tile_interior(oa), int_b = tile_interior(ob);
    const uint32_t sa = smem0 + s * stage_bytes, sb = sa + a_bytes, full = smem_u32(&full_bar[s]);
    const void *ta = recs[pi].tmap_a, *tb = recs[pi].tmap_b;
    if (lane == 0) { /* the maps live in global memory that may have held another map before */
      if (tma_a) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(ta) : "memory");
      if (tma_b) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tb) : "memory");
    }
    for (int kb = warp, use = 0; kb < nkb; kb += nst, use++) {
      const long long t0 = prof ? clock64() : 0;
      if (use >= 1) mbar_wait(smem_u32(&empty_bar[s]), (uint32_t)((use - 1) & 1));
      const long long t1 = prof ? clock64() : 0;
      if (any_tma && lane == 0) {
        asm volatile("{\n\t.reg .b64 state;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 state, [%0], %1;\n\t}" ::"r"(full), "r"(tx_bytes) : "memory");
        if (tma_a) tma_load_tile(oa, ta, sa, kb, full);
        if (tma_b) tma_load_tile(ob, tb, sb, kb, full);
      }
      long long t2 = t1, t3 = t1;
      if (any_cp) {
        if (!tma_a) load_stage(oa, int_a, sa, kb, K, lane);
        if (!tma_b) load_stage(ob, int_b, sb, kb, K, lane);
        asm volatile("cp.async.commit_group;" ::: "memory");
        t2 = prof ? clock64() : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        t3 = prof ? clock64() : 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("{\n\t.reg .b64 state;\n\tmbarrier.arrive.shared::cta.b64 state, [%0];\n\t}" ::"r"(full) : "memory");
      }
      if (prof && tid == 0) { prof[0] += t1 - t0; prof[1] += t2 - t1; prof[2] += t3 - t2; prof[3] += clock64() - t3; }
    }
  } else if (tid == RR_TC_STAGES * 32) {
    /* MMA issuer: one thread */
    for (int kb = 0; kb < nkb; kb++) {
      const int s = kb % nst;
      const long long t0 = prof ? clock64() : 0;
      mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((kb / nst) & 1));
      const long long t1 = prof ? clock64() : 0;
      if (prof) prof[4] += t1 - t0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem0 + s * stage_bytes, sb = sa + a_bytes;
#pragma unroll
      for (int j = 0; j < RR_TC_BK / 8; j++) {
        const uint64_t da = make_desc(sa + j * a_step, a_lbo, a_sbo, a_type), db = make_desc(sb + j * b_step, b_lbo, b_sbo, b_type);
        const uint32_t accumulate = (kb > 0 || j > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
            : "memory");
      }
      const long long t2 = prof ? clock64() : 0;
      /* frees the stage for its producer once these MMAs have read it */
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
      if (prof) { prof[8] += t2 - t1; prof[9] += clock64() - t2; }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
  }
  __syncwarp();
  const long long t_loop = prof ? clock64() : 0;
  mbar_wait(smem_u32(&done_bar), 0); /* every MMA has written its accumulator */
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const long long t_done = prof ? clock64() : 0;

  /* warp w reads TMEM lanes 32 (w % 4) .. + 31 (the hardware's lane window of a warp); warps 0-3 take the even 16-column
   * groups, warps 4-7 the odd ones */
  const int lane_q = warp & 3;
  const int row = tile_m * RR_TC_BM + lane_q * 32 + lane;
  const bool vec_d = (p.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.d) & 15) == 0);
  const float *auxp = p.epi == 1 ? p.aux_out : (p.epi == 2 ? p.aux_in : nullptr);
  const bool vec_aux = !auxp || ((p.ldaux % 4 == 0) && ((reinterpret_cast<uintptr_t>(auxp) & 15) == 0));
  const bool row_ok = row < p.m;
  /* the SiLU-derivative epilogue reads the pre-activations: fetch them one 16-column group ahead of the accumulator */
  float4 aux_next[4] = {};
  const int jstep = RR_TC_THREADS / 128;
  auto vec_group = [&](int j) { return row_ok && tile_n * BN + j * 16 + 16 <= n_d && vec_d && vec_aux; };
  auto fetch_aux = [&](int j) {
    const float4 *ap = reinterpret_cast<const float4 *>(p.aux_in + (size_t)row * p.ldaux + tile_n * BN + j * 16);
#pragma unroll
    for (int i = 0; i < 4; i++) aux_next[i] = ap[i];
  };
  if (p.epi == 2 && (warp >> 2) < BN / 16 && vec_group(warp >> 2)) fetch_aux(warp >> 2);
  for (int j = warp >> 2; j < BN / 16; j += jstep) {
    float4 aux_cur[4];
#pragma unroll
    for (int i = 0; i < 4; i++) aux_cur[i] = aux_next[i];
    if (p.epi == 2 && j + jstep < BN / 16 && vec_group(j + jstep)) fetch_aux(j + jstep);
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)(j * 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int c0 = tile_n * BN + j * 16;
    if (!row_ok || c0 >= n_ext) continue;
    if (vec_group(j)) {
      /* full 16-column group: 16-byte loads / stores of this thread's row; the bias comes from shared memory */
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; i++) o[i] = __uint_as_float(v[i]) + bias_s[j * 16 + i];
      if (p.epi == 1) {
        if (p.aux_out) {
          float4 *ap = reinterpret_cast<float4 *>(p.aux_out + (size_t)row * p.ldaux + c0);
#pragma unroll
          for (int i = 0; i < 4; i++) ap[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) o[i] = round_tf32(o[i] * sigmoid_fast(o[i]));
      } else if (p.epi == 2) {
        float z[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          z[4 * i] = aux_cur[i].x; z[4 * i + 1] = aux_cur[i].y; z[4 * i + 2] = aux_cur[i].z; z[4 * i + 3] = aux_cur[i].w;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float sg = sigmoid_fast(z[i]);
          o[i] = round_tf32(o[i] * sg * (1.f + z[i] * (1.f - sg)));
        }
      }
      float4 *dst = reinterpret_cast<float4 *>(p.d + (size_t)row * p.ldd + c0);
#pragma unroll
      for (int i = 0; i < 4; i++) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int c = c0 + i;
        if (c < n_d) p.d[(size_t)row * p.ldd + c] = rr_tc_epilogue(p, row, c, __uint_as_float(v[i]));
        else if (c == n_d && p.b_ones) p.ones_out[row] = __uint_as_float(v[i]);
      }
    }
  }
  if (prof && tid == RR_TC_STAGES * 32) { prof[5] = t_done - t_loop; prof[6] = t_done - t_start; prof[7] = clock64() - t_done; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
}

}  // namespace rr_tc
#endif
