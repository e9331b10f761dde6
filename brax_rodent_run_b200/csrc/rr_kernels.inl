/* rr_kernels.inl -- the fused rodent physics + run-task step, one warp per environment.
 *
 * Replaces, for one environment per warp, what the reference runs as
 *   Rodent.step (Rodent_Env_Brax.py:98-136) -> PipelineEnv.pipeline_step (:101) -> n_frames x mjx.step
 *   Rodent.reset's pipeline_init (:87) -> mjx.forward          (RR_MODE_INIT)
 *   Rodent._get_obs (:138-162) and the reward / termination terms (:106-132)
 *   brax EpisodeWrapper + AutoResetWrapper (applied inside ppo.train, brax_rodent_run_ppo.py:200)
 * Algorithms follow SURVEY.md Appendix B (mujoco-mjx 3.1.x) but in a different formulation from the
 * reference's dense one: tree-sparse mass matrix and LDL' factor (row i = ancestors of dof i), solves done
 * in registers with warp shuffles, constraint rows compacted to the active set (exact: MJX zeroes inactive
 * rows), contact Jacobians stored as 3 x chain blocks shared by the 4 pyramid rows.
 *
 * This file is included by rr_api.cu (CUDA, sm_100a).  It only uses __shfl_sync / __shfl_xor_sync /
 * __ballot_sync / __syncwarp / __ldg so that tests/emu can compile the same text for the host and run the
 * 32 lanes as fibers (test infrastructure only -- never a product path).
 */
/* This file may be included more than once with different (RR_NS, RR_WITH_DEBUG): the production kernels are compiled
 * with the debug-dump / profiling hooks removed (they double the code size and the kernel is instruction-cache bound). */
#ifndef RR_NS
#define RR_NS rr
#endif
#ifndef RR_WITH_DEBUG
#define RR_WITH_DEBUG 1
#endif

#include "../../include/rr_model_fields.h"
#include "rr_device.h"

#ifndef RR_DEV
#define RR_DEV __device__ __forceinline__
#define RR_HOSTDEV __host__ __device__ inline
#define RR_DEV_MEMBER __device__ __forceinline__
#define RR_DEV_NOINLINE __device__ __noinline__
#define RR_CLOCK() clock64()
#endif

/* model tables staged in shared memory: element `idx` of int / float table `name` */
#define RI(name, idx) (c.ti[c.m.o_##name + (idx)])
#define RF(name, idx) (c.tf[c.m.o_##name + (idx)])

/* body_anc: one byte per (round, body) */
#define RR_BODY_ANC(idx) ((int)(((unsigned)c.ti[c.m.o_body_anc + ((idx) >> 2)] >> (8 * ((idx) & 3))) & 255u))
/* packed per-entry metadata of the tree-sparse layout: row | col << 8 | rowadr[col] << 16 */
#define RR_META_ROW(x) ((x) & 255)
#define RR_META_COL(x) (((x) >> 8) & 255)
#define RR_META_ANCRADR(x) ((int)((unsigned)(x) >> 16))

/* CTA-wide rendezvous used only to keep the warps of a CTA in the same code region (instruction-cache locality);
 * it carries no data dependence.  No-op in the host emulator. */
/* Live warps (substep) and padding warps (substep_idle) reach the rendezvous from different call sites under a condition
 * that varies per warp.  PTX defines `bar.sync 0` by arrival COUNT, wherever the arriving warps are in the program (it only
 * has to be warp-uniform, which it is: the condition is per warp); the C++ __syncthreads() rule about divergent code does
 * not cover that, so the barrier is written as the PTX instruction itself. */
#ifndef RR_CTA_SYNC
#define RR_CTA_SYNC() asm volatile("bar.sync 0;" ::: "memory")
#endif
#ifndef RR_DUP_SOLVE
#define RR_DUP_SOLVE 0
#endif
#ifndef RR_DUP_MULJ
#define RR_DUP_MULJ 0
#endif
#ifndef RR_DUP_MULJT
#define RR_DUP_MULJT 0
#endif
#ifndef RR_DUP_LS
#define RR_DUP_LS 0
#endif
#ifndef RR_DUP_KIN
#define RR_DUP_KIN 0
#endif
#ifndef RR_DUP_RNE
#define RR_DUP_RNE 0
#endif
#ifndef RR_DUP_CRB
#define RR_DUP_CRB 0
#endif
#ifndef RR_DUP_FACTOR
#define RR_DUP_FACTOR 0
#endif
#ifndef RR_MV_RECURRENCE
#define RR_MV_RECURRENCE 1
#endif
#ifndef RR_SYNC_LEVEL
#define RR_SYNC_LEVEL 2 /* 1: per substep; 2: + before the factorisations (and the collision phase if RR_SYNC_COLLIDE);
                           3: + per CG iteration; 4: + between the line search and the gradient update of an iteration.
                           With the v15 code size (17 k instructions, I-cache hit rate 93 %) the per-substep rendezvous is
                           what matters (none: 2.1x slower; every 2nd substep: -22 %); levels 1 / 2 / 3 / 4 run within 1 %. */
#endif
#ifndef RR_SYNC_PERIOD
#define RR_SYNC_PERIOD 1 /* substep-start rendezvous every this many substeps */
#endif
#ifndef RR_SYNC_FACTOR
#define RR_SYNC_FACTOR 1
#endif
#ifndef RR_SYNC_COLLIDE
#define RR_SYNC_COLLIDE 1 /* + before the collision / constraint phase: 637 k -> 642 k env-steps/s (round 2) */
#endif
#ifndef RR_SYNC_EULER
#define RR_SYNC_EULER 0
#endif
#ifndef RR_SYNC_RNE
#define RR_SYNC_RNE 0
#endif
#define RR_CTA_SYNC_IF(flag) do { if (flag) RR_CTA_SYNC(); } while (0)
#define RR_CTA_SYNC_AT(level)                \
  do {                                       \
    if (RR_SYNC_LEVEL >= (level)) RR_CTA_SYNC(); \
  } while (0)

/* 1: right-looking (scatter) L'DL (factor2_rl), 0: gather form (factor2).  Measured: 415 k vs 571 k env-steps/s -- the scatter
 * form moves every Schur-complement entry through shared memory twice per descendant; kept for the record only. */
#ifndef RR_FACTOR_RL
#define RR_FACTOR_RL 0
#endif
#define RR_PRAGMA_(x) _Pragma(#x)
#define RR_UNROLL(n) RR_PRAGMA_(unroll n)

/* Raw shared-memory addressing for the solve loops: a 32-bit shared-window byte address per lane plus a byte offset is
 * ONE integer add in front of the load (generic C++ indexing costs an add and a shift-add).  The host emulator provides
 * plain-pointer versions. */
#ifndef RR_SADDR_T
#define RR_SADDR_T unsigned
#define RR_SADDR(p) ((unsigned)__cvta_generic_to_shared(p))
__device__ __forceinline__ float rr_lds_f32(unsigned a) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
#define RR_SLOAD(a) rr_lds_f32(a)
/* (an inline-PTX predicated load + in-place predicated FMA was tried for this update: ptxas then spills 4 KB) */
#define RR_SOLVE_UPDATE(x, m, bit, a, xi) do { if ((m) & (bit)) (x) -= RR_SLOAD(a) * (xi); } while (0)
#endif

#define RR_FULL 0xffffffffu
#define RR_MINVAL 1e-15f
#define RR_MINIMP 0.0001f
#define RR_MAXIMP 0.9999f
#define RR_SIGN_BIT 0x40000000

/* Reciprocal on the critical paths (pivot of a factorisation row, Newton steps of the line search): the IEEE division is
 * a 7-instruction sequence with a branch to a slow path (ptxas cannot schedule across it); MUFU.RCP plus one Newton step is
 * 3 straight-line instructions and accurate to about 1 ulp for the normal, positive arguments met here. */
#ifndef RR_RCP
__device__ __forceinline__ float rr_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);
}
#define RR_RCP(x) rr_rcp(x)
/* sqrt without the IEEE fix-up / slow path (1 ulp) */
__device__ __forceinline__ float rr_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#define RR_SQRT(x) rr_sqrt(x)
/* sin and cos of a joint half-angle: quadrant reduction (two-constant Cody-Waite, exact enough far beyond any joint range)
 * and the classic single-precision minimax polynomials on [-pi/4, pi/4] (about 1 ulp), straight-line -- libdevice's
 * sincosf carries a branch to a Payne-Hanek slow path inside the serial tree walk of the kinematics */
__device__ __forceinline__ void rr_sincos(float x, float *sn, float *cs) {
  const float q = rintf(x * 0.636619772367581343f);
  float r = fmaf(q, -1.5707962512969971f, x);
  r = fmaf(q, -7.5497894158615964e-8f, r);
  const float r2 = r * r;
  float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, r2, -1.6666654611e-1f);
  ps = fmaf(ps * r2, r, r);
  float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, r2, 4.166664568298827e-2f);
  pc = fmaf(pc * r2, r2, fmaf(r2, -0.5f, 1.f));
  const int k = (int)q;
  const float s0 = (k & 1) ? pc : ps, c0 = (k & 1) ? ps : pc;
  *sn = (k & 2) ? -s0 : s0;
  *cs = ((k + 1) & 2) ? -c0 : c0;
}
#define RR_SINCOS(x, s, c) rr_sincos(x, s, c)
#endif

namespace RR_NS {

RR_DEV float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(RR_FULL, x, o);
  return x;
}

/* ------------------------------------------------------------------------------------------ small math */
RR_DEV void quat_mul(float *r, const float *a, const float *b) {
  float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
RR_DEV void cross3(float *r, const float *a, const float *b) {
  float x = a[1] * b[2] - a[2] * b[1];
  float y = a[2] * b[0] - a[0] * b[2];
  float z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
RR_DEV float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* mjx math.rotate: r = 2 (u.v) u + (s^2 - u.u) v + 2 s (u x v) */
RR_DEV void rotq(float *r, const float *v, const float *q) {
  float s = q[0];
  const float *u = q + 1;
  float uv = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(c, u, v);
  float k = s * s - uu;
#pragma unroll
  for (int i = 0; i < 3; i++) r[i] = 2.f * (uv * u[i]) + k * v[i] + 2.f * s * c[i];
}
RR_DEV void quat_to_mat(float *m, const float *q) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2.f * (x * y - w * z); m[2] = 2.f * (x * z + w * y);
  m[3] = 2.f * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2.f * (y * z - w * x);
  m[6] = 2.f * (x * z - w * y); m[7] = 2.f * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
RR_DEV void axis_angle_quat(float *q, const float *axis, float angle) {
  float s, c;
  RR_SINCOS(angle * 0.5f, &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
RR_DEV float normalize3(float *v) {
  float n = RR_SQRT(dot3(v, v));
  float d = n + 1e-6f * (n == 0.f ? 1.f : 0.f);
  const float id = RR_RCP(d);
  v[0] *= id; v[1] *= id; v[2] *= id;
  return n;
}
RR_DEV void normalize4(float *v) {
  float n = RR_SQRT(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
  float d = n + 1e-6f * (n == 0.f ? 1.f : 0.f);
  const float id = RR_RCP(d);
  v[0] *= id; v[1] *= id; v[2] *= id; v[3] *= id;
}
/* mjx math.inert_mul: cinert(10) x motion(6: ang, lin) -> force(6) */
RR_DEV void inert_mul(float *r, const float *i, const float *v) {
  float c1[3], c2[3];
  cross3(c1, i + 6, v + 3);
  cross3(c2, i + 6, v);
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + c1[0];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + c1[1];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + c1[2];
  r[3] = i[9] * v[3] - c2[0];
  r[4] = i[9] * v[4] - c2[1];
  r[5] = i[9] * v[5] - c2[2];
}
RR_DEV void motion_cross(float *r, const float *u, const float *v) {
  float a[3], b[3], c[3];
  cross3(a, u, v);
  cross3(b, u + 3, v);
  cross3(c, u, v + 3);
#pragma unroll
  for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
RR_DEV void motion_cross_force(float *r, const float *v, const float *f) {
  float a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
#pragma unroll
  for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}
RR_DEV float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* constraint.py _kbi: stiffness k, damping b and impedance for one row (Appendix B.5) */
RR_DEV_NOINLINE void kbi(float timestep, const float *solref, const float *solimp, float pos, float &k, float &b, float &imp) {
  float timeconst = fmaxf(solref[0], 2.f * timestep), dampratio = solref[1];
  float dmin = clampf(solimp[0], RR_MINIMP, RR_MAXIMP), dmax = clampf(solimp[1], RR_MINIMP, RR_MAXIMP);
  float width = fmaxf(solimp[2], RR_MINVAL), mid = clampf(solimp[3], RR_MINIMP, RR_MAXIMP), power = fmaxf(solimp[4], 1.f);
  k = RR_RCP(dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = 2.f * RR_RCP(dmax * timeconst);
  if (solref[0] <= 0.f) k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0.f) b = -solref[1] / dmax;
  float x = fabsf(pos) * RR_RCP(width);
  float ia, ib;
  if (power == 2.f) { /* the MuJoCo default (every rodent model): no powf */
    ia = RR_RCP(mid) * (x * x);
    ib = 1.f - RR_RCP(1.f - mid) * ((1.f - x) * (1.f - x));
  } else {
    ia = (1.f / powf(mid, power - 1.f)) * powf(x, power);
    ib = 1.f - (1.f / powf(1.f - mid, power - 1.f)) * powf(1.f - x, power);
  }
  float y = x < mid ? ia : ib;
  float im = clampf(dmin + y * (dmax - dmin), dmin, dmax);
  imp = x > 1.f ? dmax : im;
}

/* ------------------------------------------------------------------------------------------ context */
template <int NS>
struct Ctx {
  const RRModelDev &m;
  const RRStepArgs &a;
  int env, lane;
  const int32_t *ti; /* shared-memory copies of the model tables */
  const float *tf;
  float *qpos, *qvel, *act, *ctrl, *actdot, *xpos, *xquat, *com, *cinert, *cdof, *cvel, *M, *LD, *vbuf, *qfrc_act;
  float *crb, *fcrb, *cacc, *cfrc, *con_dist, *cab, *cscr, *cbv, *row_D, *row_aref, *row_Jaref, *row_jv;
  int *row_id, *cact, *ckidx;
  float *prof_acc; /* RR_NPROF per-stage cycle sums (debug builds) */
  float *xq1; /* xquat of body 1 saved for the observation (xquat itself is recycled by the solver phase) */
  float *sm_base, *grows; /* shared-memory base; global overflow for constraint rows */
  /* per-lane dof metadata: dof i = lane + 32 s */
  int radr[NS], dep[NS], nd[NS];
  float dinv[NS], dinv2[NS]; /* 1 / D of LD (M) and of the Euler factor (M + dt damping) */
  float ma_warm[NS];         /* M qacc_warmstart, formed before M is factorised in place */
  /* register-resident nv-vectors */
  float qfrc_smooth[NS], qacc_smooth[NS], warm[NS], qacc[NS], qfrc_constraint[NS];
  int nla, nca; /* active limit rows, active contacts; rows = nla + 4 nca */
  bool live;         /* false: padding pass of a persistent warp (keeps CTA barriers matched); no global stores */
  bool last_substep; /* the forward pass whose cinert / cvel / qfrc_actuator the observation reports */
  int niter;
  int niter_total; /* CG iterations over all substeps of this step: the load-balancing cost estimate */
  float *dbg;
  long long tprev;

  /* constraint rows live in shared memory when at most capR are active, else in the per-warp global scratch */
  RR_DEV_MEMBER void use_rows(bool in_smem) {
    const RRSmem &s = m.sm;
    float *base = in_smem ? sm_base + s.row_D : grows;
    int stride = in_smem ? s.capR : m.nefc;
    row_D = base; row_aref = base + stride; row_Jaref = base + 2 * stride; row_jv = base + 3 * stride;
    row_id = (int *)(base + 4 * stride);
  }

  RR_DEV_MEMBER Ctx(const RRModelDev &m_, const RRStepArgs &a_, int env_, int slot_, float *sm, const int32_t *ti_, const float *tf_,
                    int lane_)
      : m(m_), a(a_), env(env_), lane(lane_), ti(ti_), tf(tf_) {
    const RRSmem &s = m.sm;
    sm_base = sm;
    grows = a.scratch + (size_t)slot_ * a.scratch_stride;
    xq1 = sm + s.xq1;
    prof_acc = grows + a.scratch_stride - 32; /* instrumented builds only: the tail of the per-warp global scratch */
    qpos = sm + s.qpos; qvel = sm + s.qvel; act = sm + s.act; ctrl = sm + s.ctrl; actdot = sm + s.actdot;
    xpos = sm + s.xpos; xquat = sm + s.xquat; com = sm + s.com; cinert = sm + s.cinert; cdof = sm + s.cdof;
    cvel = sm + s.cvel; M = sm + s.M; LD = sm + s.LD;  vbuf = sm + s.vbuf; qfrc_act = sm + s.qfrc_act;
    crb = sm + s.crb; fcrb = sm + s.fcrb; cacc = sm + s.cacc; cfrc = sm + s.cfrc;
    con_dist = sm + s.con_dist; cab = sm + s.cab; cscr = sm + s.cscr; cbv = sm + s.cbv;
    cact = (int *)(sm + s.cact);
    ckidx = (int *)(sm + s.ckidx);
    use_rows(true);
#pragma unroll
    for (int s_ = 0; s_ < NS; s_++) {
      int i = lane + 32 * s_;
      bool v = i < m.nv;
      radr[s_] = v ? ti[m.o_dof_rowadr + i] : 0;
      dep[s_] = v ? ti[m.o_dof_depth + i] : 0;
      nd[s_] = v ? ti[m.o_dof_ndesc + i] : 0;
      dinv[s_] = 0.f; dinv2[s_] = 0.f; ma_warm[s_] = 0.f;
    }
    nla = nca = 0;
    niter = 0;
    last_substep = false;
    live = true;
    niter_total = 0;
    dbg = a.dbg.buf ? a.dbg.buf + (size_t)env * a.dbg.stride : nullptr;
    tprev = 0;
  }
};

#define RR_FOR_S _Pragma("unroll") for (int s = 0; s < NS; s++)

/* per-stage cycle counters accumulate in shared memory (a global read-modify-write per call would dominate the short
 * stages) and are flushed once per environment by prof_flush */
template <int NS>
RR_DEV void prof(Ctx<NS> &c, int id) {
  if (RR_WITH_DEBUG && c.a.prof && c.live) {
    long long t = RR_CLOCK();
    if (c.lane == 0) c.prof_acc[id] += (float)(t - c.tprev);
    c.tprev = RR_CLOCK();
  }
}
template <int NS>
RR_DEV void prof_flush(Ctx<NS> &c) {
  if (RR_WITH_DEBUG && c.a.prof && c.live && c.lane == 0)
    for (int i = 0; i < RR_NPROF; i++) c.a.prof[(size_t)c.env * RR_NPROF + i] += (long long)c.prof_acc[i];
}
template <int NS>
RR_DEV float vdot(const float (&x)[NS], const float (&y)[NS]) {
  float t = 0.f;
  RR_FOR_S t += x[s] * y[s];
  return warp_sum(t);
}
template <int NS>
RR_DEV void vload(const Ctx<NS> &c, float (&x)[NS], const float *buf) {
  RR_FOR_S { int i = c.lane + 32 * s; x[s] = i < c.m.nv ? buf[i] : 0.f; }
}
template <int NS>
RR_DEV void vstore(const Ctx<NS> &c, const float (&x)[NS], float *buf) {
  RR_FOR_S { int i = c.lane + 32 * s; if (i < c.m.nv) buf[i] = x[s]; }
}
template <int NS>
RR_DEV float vselect(const float (&x)[NS], int slot) {
  float r = x[0];
  RR_FOR_S if (s == slot) r = x[s];
  return r;
}

/* debug dump helpers (parity tests only; no effect when dbg == nullptr) */
enum {
  RR_DBG_XPOS = 0, RR_DBG_XQUAT, RR_DBG_COM, RR_DBG_CINERT, RR_DBG_CDOF, RR_DBG_CVEL, RR_DBG_M, RR_DBG_LD,
  RR_DBG_QFRC_BIAS, RR_DBG_QFRC_PASSIVE, RR_DBG_QFRC_ACTUATOR, RR_DBG_QFRC_SMOOTH, RR_DBG_QACC_SMOOTH,
  RR_DBG_CON_DIST, RR_DBG_CON_POS, RR_DBG_CON_FRAME, RR_DBG_EFC_J, RR_DBG_EFC_D, RR_DBG_EFC_AREF, RR_DBG_EFC_FORCE,
  RR_DBG_QACC, RR_DBG_QFRC_CONSTRAINT, RR_DBG_SCALARS, RR_DBG_NFIELDS
};
RR_HOSTDEV int dbg_count(const RRModelDev &m, int f) {
  switch (f) {
    case RR_DBG_XPOS: return 3 * m.nbody;
    case RR_DBG_XQUAT: return 4 * m.nbody;
    case RR_DBG_COM: return 3 * m.nroot;
    case RR_DBG_CINERT: return 10 * m.nbody;
    case RR_DBG_CDOF: return 6 * m.nv;
    case RR_DBG_CVEL: return 6 * m.nbody;
    case RR_DBG_M: case RR_DBG_LD: return m.nM;
    case RR_DBG_QFRC_BIAS: case RR_DBG_QFRC_PASSIVE: case RR_DBG_QFRC_ACTUATOR: case RR_DBG_QFRC_SMOOTH:
    case RR_DBG_QACC_SMOOTH: case RR_DBG_QACC: case RR_DBG_QFRC_CONSTRAINT: return m.nv;
    case RR_DBG_CON_DIST: return m.ncon;
    case RR_DBG_CON_POS: return 3 * m.ncon;
    case RR_DBG_CON_FRAME: return 9 * m.ncon;
    case RR_DBG_EFC_J: return m.nefc * m.nv;
    case RR_DBG_EFC_D: case RR_DBG_EFC_AREF: case RR_DBG_EFC_FORCE: return m.nefc;
    case RR_DBG_SCALARS: return 8; /* niter, nla, nca, cost, ... */
  }
  return 0;
}
RR_HOSTDEV int dbg_offset(const RRModelDev &m, int f) {
  int o = 0;
  for (int k = 0; k < f; k++) o += dbg_count(m, k);
  return o;
}
template <int NS>
RR_DEV void dbg_copy(Ctx<NS> &c, int field, const float *src, int n) {
  if (!RR_WITH_DEBUG || !c.dbg) return;
  float *dst = c.dbg + dbg_offset(c.m, field);
  for (int i = c.lane; i < n; i += 32) dst[i] = src[i];
}
template <int NS>
RR_DEV void dbg_vec(Ctx<NS> &c, int field, const float (&x)[NS]) {
  if (!RR_WITH_DEBUG || !c.dbg) return;
  float *dst = c.dbg + dbg_offset(c.m, field);
  RR_FOR_S { int i = c.lane + 32 * s; if (i < c.m.nv) dst[i] = x[s]; }
}

/* ------------------------------------------------------------------------------------------ kinematics (B.1) */
/* Tree scan by pointer doubling instead of a walk over the tree levels (27 dependent levels of quaternion algebra for the
 * rodent, with 1 - 5 busy lanes each): every body first forms its transform RELATIVE to its effective parent (joint
 * rotations included; all bodies in parallel), then nround = ceil(log2(depth)) rounds compose T_b <- T_anc_k(b) o T_b with
 * anc_k = the 2^k-th ancestor (body_anc), all bodies per round in parallel, double-buffered between (xpos, xquat) and the
 * LD array (dead until the factorisation).  World poses are identical to the level walk up to the association order of the
 * products (1e-7 relative).  Joint anchors / axes are formed in the parent frame and mapped to the world afterwards. */
template <int NS>
RR_DEV void kinematics(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  const int nb = m.nbody, R = m.nround;
  /* buffer 0 = (xpos, xquat); buffer 1 = LD as 8 floats per body (pos, -, quat).  Start so that round R ends in buffer 0. */
  float *p_in = (R & 1) ? c.LD : c.xpos, *q_in = (R & 1) ? c.LD + 4 : c.xquat;
  int sp_in = (R & 1) ? 8 : 3, sq_in = (R & 1) ? 8 : 4;
  float *p_out = (R & 1) ? c.xpos : c.LD, *q_out = (R & 1) ? c.xquat : c.LD + 4;
  int sp_out = (R & 1) ? 3 : 8, sq_out = (R & 1) ? 4 : 8;
  __syncwarp();
  for (int b = c.lane; b < nb; b += 32) {
    float pos[3], quat[4], r[3];
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = RF(body_epos, 3 * b + k);
#pragma unroll
    for (int k = 0; k < 4; k++) quat[k] = RF(body_equat, 4 * b + k);
    const int jadr = RI(body_jntadr, b), jnum = RI(body_jntnum, b);
    for (int j = jadr; j < jadr + jnum; j++) {
      const int qa = RI(jnt_qposadr, j), da = RI(jnt_dofadr, j);
      if (RI(jnt_type, j) == RR_JNT_FREE) {
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] = c.qpos[qa + k];
#pragma unroll
        for (int k = 0; k < 4; k++) quat[k] = c.qpos[qa + 3 + k];
        normalize4(quat);
#pragma unroll
        for (int k = 0; k < 4; k++) c.qpos[qa + 3 + k] = quat[k]; /* normalised quaternion is written back */
      } else {
        float jp[3], ja[3], anchor[3], axis[3], qloc[4], q2[4];
#pragma unroll
        for (int k = 0; k < 3; k++) { jp[k] = RF(jnt_pos, 3 * j + k); ja[k] = RF(jnt_axis, 3 * j + k); }
        rotq(r, jp, quat);
#pragma unroll
        for (int k = 0; k < 3; k++) anchor[k] = r[k] + pos[k];
        rotq(axis, ja, quat);
        axis_angle_quat(qloc, ja, c.qpos[qa] - RF(qpos0, qa));
        quat_mul(q2, quat, qloc);
#pragma unroll
        for (int k = 0; k < 4; k++) quat[k] = q2[k];
        rotq(r, jp, quat);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          pos[k] = anchor[k] - r[k];
          c.cdof[6 * da + k] = axis[k];       /* temp: joint axis, parent frame */
          c.cdof[6 * da + 3 + k] = anchor[k]; /* temp: joint anchor, parent frame */
        }
      }
    }
    if (b == 0) { pos[0] = pos[1] = pos[2] = 0.f; quat[0] = 1.f; quat[1] = quat[2] = quat[3] = 0.f; }
#pragma unroll
    for (int k = 0; k < 3; k++) p_in[sp_in * b + k] = pos[k];
#pragma unroll
    for (int k = 0; k < 4; k++) q_in[sq_in * b + k] = quat[k];
  }
  __syncwarp();
#pragma unroll 1
  for (int rd = 0; rd < R; rd++) {
    for (int b = c.lane; b < nb; b += 32) {
      const int an = RR_BODY_ANC(rd * nb + b);
      float ap[3], aq[4], bp[3], bq[4], r[3], q2[4];
#pragma unroll
      for (int k = 0; k < 3; k++) { ap[k] = p_in[sp_in * an + k]; bp[k] = p_in[sp_in * b + k]; }
#pragma unroll
      for (int k = 0; k < 4; k++) { aq[k] = q_in[sq_in * an + k]; bq[k] = q_in[sq_in * b + k]; }
      rotq(r, bp, aq);
      quat_mul(q2, aq, bq);
#pragma unroll
      for (int k = 0; k < 3; k++) p_out[sp_out * b + k] = ap[k] + r[k];
#pragma unroll
      for (int k = 0; k < 4; k++) q_out[sq_out * b + k] = q2[k];
    }
    __syncwarp();
    { float *t = p_in; p_in = p_out; p_out = t; t = q_in; q_in = q_out; q_out = t; }
    { int t = sp_in; sp_in = sp_out; sp_out = t; t = sq_in; sq_in = sq_out; sq_out = t; }
  }
  /* world poses are now in (xpos, xquat): inertial-frame origins, then joint anchors / axes to the world */
  for (int b = c.lane; b < nb; b += 32) {
    float pos[3], quat[4], ip[3], r[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { pos[k] = c.xpos[3 * b + k]; ip[k] = RF(body_ipos, 3 * b + k); }
#pragma unroll
    for (int k = 0; k < 4; k++) quat[k] = c.xquat[4 * b + k];
    rotq(r, ip, quat);
#pragma unroll
    for (int k = 0; k < 3; k++) c.cinert[10 * b + 6 + k] = pos[k] + r[k]; /* temp: xipos */
  }
  for (int j = c.lane; j < m.njnt; j += 32) {
    if (RI(jnt_type, j) == RR_JNT_FREE) continue;
    const int da = RI(jnt_dofadr, j), p = RI(body_eparent, RI(jnt_bodyid, j));
    float pp[3], pq[4], ax[3], an[3], r1[3], r2[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { pp[k] = c.xpos[3 * p + k]; ax[k] = c.cdof[6 * da + k]; an[k] = c.cdof[6 * da + 3 + k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) pq[k] = c.xquat[4 * p + k];
    rotq(r1, ax, pq);
    rotq(r2, an, pq);
#pragma unroll
    for (int k = 0; k < 3; k++) { c.cdof[6 * da + k] = r1[k]; c.cdof[6 * da + 3 + k] = pp[k] + r2[k]; } /* temp: xaxis, xanchor */
  }
  __syncwarp();
}

/* ------------------------------------------------------------------------------------------ com_pos (B.2) */
template <int NS>
RR_DEV void com_pos(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  for (int r = 0; r < m.nroot; r++) {
    float sx = 0.f, sy = 0.f, sz = 0.f, sm = 0.f;
    for (int b = 1 + c.lane; b < m.nbody; b += 32) {
      if (RI(body_rootslot, b) == r) {
        float mass = RF(body_mass, b);
        sx += c.cinert[10 * b + 6] * mass; sy += c.cinert[10 * b + 7] * mass; sz += c.cinert[10 * b + 8] * mass;
        sm += mass;
      }
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sm = warp_sum(sm);
    if (c.lane == 0) {
      bool tiny = sm < RR_MINVAL;
      const float ism = RR_RCP(tiny ? 1.f : sm);
      c.com[3 * r + 0] = tiny ? 0.f : sx * ism;
      c.com[3 * r + 1] = tiny ? 0.f : sy * ism;
      c.com[3 * r + 2] = tiny ? 0.f : sz * ism;
    }
  }
  __syncwarp();
  /* cinert: inertia about the tree COM, (Ixx Iyy Izz Ixy Ixz Iyz, m*off, m) */
  for (int b = c.lane; b < m.nbody; b += 32) {
    float *ci = c.cinert + 10 * b;
    if (b == 0) {
#pragma unroll
      for (int k = 0; k < 10; k++) ci[k] = 0.f;
      continue;
    }
    float q[4], iq[4], xq[4], R[9], I[3], off[3];
#pragma unroll
    for (int k = 0; k < 4; k++) { xq[k] = c.xquat[4 * b + k]; iq[k] = RF(body_iquat, 4 * b + k); }
    quat_mul(q, xq, iq);
    quat_to_mat(R, q);
    int rs = RI(body_rootslot, b);
    float mb = RF(body_mass, b);
#pragma unroll
    for (int k = 0; k < 3; k++) { I[k] = RF(body_inertia, 3 * b + k); off[k] = ci[6 + k] - c.com[3 * rs + k]; }
    float d2 = dot3(off, off);
    float A00 = R[0] * I[0] * R[0] + R[1] * I[1] * R[1] + R[2] * I[2] * R[2] + mb * (d2 - off[0] * off[0]);
    float A11 = R[3] * I[0] * R[3] + R[4] * I[1] * R[4] + R[5] * I[2] * R[5] + mb * (d2 - off[1] * off[1]);
    float A22 = R[6] * I[0] * R[6] + R[7] * I[1] * R[7] + R[8] * I[2] * R[8] + mb * (d2 - off[2] * off[2]);
    float A01 = R[0] * I[0] * R[3] + R[1] * I[1] * R[4] + R[2] * I[2] * R[5] - mb * off[0] * off[1];
    float A02 = R[0] * I[0] * R[6] + R[1] * I[1] * R[7] + R[2] * I[2] * R[8] - mb * off[0] * off[2];
    float A12 = R[3] * I[0] * R[6] + R[4] * I[1] * R[7] + R[5] * I[2] * R[8] - mb * off[1] * off[2];
    ci[0] = A00; ci[1] = A11; ci[2] = A22; ci[3] = A01; ci[4] = A02; ci[5] = A12;
    ci[6] = off[0] * mb; ci[7] = off[1] * mb; ci[8] = off[2] * mb; ci[9] = mb;
  }
  /* cdof: [ang; lin] about the tree COM */
  for (int j = c.lane; j < m.njnt; j += 32) {
    int b = RI(jnt_bodyid, j), da = RI(jnt_dofadr, j);
    int rs = RI(body_rootslot, b);
    float *cd = c.cdof + 6 * da;
    if (RI(jnt_type, j) == RR_JNT_FREE) {
      float R[9], xq[4], off[3];
#pragma unroll
      for (int k = 0; k < 4; k++) xq[k] = c.xquat[4 * b + k];
      quat_to_mat(R, xq);
#pragma unroll
      for (int k = 0; k < 3; k++) off[k] = c.com[3 * rs + k] - c.xpos[3 * b + k];
#pragma unroll
      for (int d = 0; d < 3; d++) {
#pragma unroll
        for (int k = 0; k < 6; k++) cd[6 * d + k] = (k == 3 + d) ? 1.f : 0.f;
        float ax[3] = {R[d], R[3 + d], R[6 + d]}, cr[3];
        cross3(cr, ax, off);
#pragma unroll
        for (int k = 0; k < 3; k++) { cd[6 * (3 + d) + k] = ax[k]; cd[6 * (3 + d) + 3 + k] = cr[k]; }
      }
    } else {
      float ax[3], off[3], cr[3];
#pragma unroll
      for (int k = 0; k < 3; k++) { ax[k] = cd[k]; off[k] = c.com[3 * rs + k] - cd[3 + k]; }
      cross3(cr, ax, off);
#pragma unroll
      for (int k = 0; k < 3; k++) cd[3 + k] = cr[k];
    }
  }
  __syncwarp();
}

/* ------------------------------------------------------------------------------------------ crb + qM (B.3) */
template <int NS>
RR_DEV void crb_and_mass_matrix(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  for (int i = c.lane; i < 10 * m.nbody; i += 32) c.crb[i] = c.cinert[i];
  __syncwarp();
  /* lane q owns component q of every body: no cross-lane hazard, so no rendezvous inside the loop */
#pragma unroll 4
  for (int b = m.nbody - 1; b > 0; b--) {
    const int p = m.kpar[b];
    if (c.lane < 10 && p > 0) c.crb[10 * p + c.lane] += c.crb[10 * b + c.lane];
  }
  __syncwarp();
  /* row i of qM: f = crb[body(i)] * cdof[i], then M(i, a) = cdof[a] . f for every ancestor-or-self a of i */
  RR_FOR_S {
    int i = c.lane + 32 * s;
    if (i < m.nv) {
      float f[6], cr[10], cd[6];
      int b = RI(dof_bodyid, i);
#pragma unroll
      for (int k = 0; k < 10; k++) cr[k] = c.crb[10 * b + k];
#pragma unroll
      for (int k = 0; k < 6; k++) cd[k] = c.cdof[6 * i + k];
      inert_mul(f, cr, cd);
      const int adr = c.radr[s], dp = c.dep[s];
      for (int t = 0; t <= dp; t++) {
        const float *cj = c.cdof + 6 * RR_META_COL(RI(M_meta, adr + t));
        float v = cj[0] * f[0] + cj[1] * f[1] + cj[2] * f[2] + cj[3] * f[3] + cj[4] * f[4] + cj[5] * f[5];
        if (t == dp) v += RF(dof_armature, i);
        c.M[adr + t] = v;
      }
    }
  }
  __syncwarp();
}

/* Tree-sparse L'DL factorisations (the factorisation MuJoCo's mj_factorM computes, leaves to root; equivalent to the
 * dense Cholesky MJX runs, jax.scipy cho_factor, up to rounding) of BOTH matrices a substep needs, in one sweep:
 *   LD  <- factor(M)                        (solver: M^-1 grad, qacc_smooth)
 *   M   <- factor(M + dt diag(damping))     (implicit-damping Euler), in place over M, which is dead afterwards
 * The two eliminations are independent, so interleaving them doubles the instruction-level parallelism of the
 * latency-bound sweep (a lone warp spends ~1.1 k cycles per row either way).
 * Gather form: row k is final once the rows of all its descendants j are,
 *   row_k[s] = M(k, a_s) - sum_j L(j, k) D_j L(j, a_s),   D_k = row_k[diag],   L(k, a_s) = row_k[s] / D_k,
 * lane s accumulates entry s in a register (rows deeper than 32 take a second register); the per-descendant scalars
 * w_j = L(j, k) D_j are staged once per row.  On exit the diagonal slots hold 1 / D. */
template <int NS>
RR_DEV void factor2(Ctx<NS> &c, float dt) {
  const RRModelDev &m = c.m;
  /* staging, per descendant j of the step's rows: w = L(j, row) D_j for up to 4 rows x 2 matrices (two float4) and the row
   * address of j.  The C1 region (cinert .. cfrc) is dead here: its observation slices were stored by forward_outputs. */
  const int nd4 = (m.nv + 3) & ~3;
  float4 *stA = reinterpret_cast<float4 *>(c.cinert), *stB = stA + nd4;
  int *stR = reinterpret_cast<int *>(c.cinert + 8 * nd4);
  float *LD = c.LD, *L2 = c.M;
  __syncwarp();
#pragma unroll 1
  for (int k = m.nv - 1; k >= 0;) {
    const int pk = RI(dof_pack, k); /* rowadr | depth << 16 | ndesc << 24 */
    const int adr = pk & 0xffff, mk = (pk >> 16) & 255, nd = (int)((unsigned)pk >> 24);
    const int lwp = RI(dof_log2w, k); /* bits 0-7: log2 of the row width rounded up; bits 8-9: rows k-1 .. k-T+1 ride along */
    if (mk < 32) {
      /* Up to four rows per step.  k-1 .. k-T+1 are the successive parents of k along a chain without other children, so
       * their descendants are those of k plus the rows of the step itself: all rows run over the same descendant rows
       * (loaded once) and the contribution of a finished row to the remaining ones is applied from registers.  Short rows
       * leave lanes idle: the descendants are split over G = 32 / W lane groups (W = width of row k rounded up to a power
       * of two) and the partial sums added with shuffles. */
      const int T = ((lwp >> 8) & 3) + 1;
      const int lw = lwp & 255, W = 1 << lw, G = 32 >> lw;
      const int s0 = c.lane & (W - 1), g = c.lane >> lw;
      const bool on = s0 <= mk;
      const int so = on ? s0 : 0;
      int adr_r[4];
      float acc[4][2];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        /* rows are stored back to back and row k - r has mk - r + 1 entries: rowadr[k - r] = adr - r (mk + 1) + r (r + 1) / 2 */
        adr_r[r] = adr - r * (mk + 1) + (r * (r + 1)) / 2;
        const bool mine = r < T && g == 0 && s0 <= mk - r;
        const float v = mine ? L2[adr_r[r] + s0] : 0.f;
        acc[r][0] = v;
        acc[r][1] = v; /* dt damping joins the diagonal when the row is finished (one uniform add) */
      }
      for (int jj = c.lane; jj < nd; jj += 32) {
        const int pj = RI(dof_pack, k + 1 + jj), rj = pj & 0xffff, dj = (pj >> 16) & 255;
        const float D1 = LD[rj + dj], D2 = L2[rj + dj];
        float w[4][2];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int e = rj + (r < T ? mk - r : mk);
          w[r][0] = r < T ? LD[e] * D1 : 0.f;
          w[r][1] = r < T ? L2[e] * D2 : 0.f;
        }
        stA[jj] = make_float4(w[0][0], w[0][1], w[1][0], w[1][1]);
        if (T > 2) stB[jj] = make_float4(w[2][0], w[2][1], w[3][0], w[3][1]);
        stR[jj] = rj;
      }
      __syncwarp();
      if (T > 2) {
#pragma unroll 2
        for (int jj = g; jj < nd; jj += G) {
          const float4 wa = stA[jj], wb = stB[jj];
          const int rj = stR[jj];
          const float l1 = LD[rj + so], l2 = L2[rj + so]; /* lanes beyond the row width accumulate values nobody reads */
          acc[0][0] -= wa.x * l1; acc[0][1] -= wa.y * l2; acc[1][0] -= wa.z * l1; acc[1][1] -= wa.w * l2;
          acc[2][0] -= wb.x * l1; acc[2][1] -= wb.y * l2; acc[3][0] -= wb.z * l1; acc[3][1] -= wb.w * l2;
        }
      } else {
#pragma unroll 4
        for (int jj = g; jj < nd; jj += G) {
          const float4 wa = stA[jj];
          const int rj = stR[jj];
          const float l1 = LD[rj + so], l2 = L2[rj + so];
          acc[0][0] -= wa.x * l1; acc[0][1] -= wa.y * l2; acc[1][0] -= wa.z * l1; acc[1][1] -= wa.w * l2;
        }
      }
      for (int o = W; o < 32; o <<= 1) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          if (r < T) {
            acc[r][0] += __shfl_xor_sync(RR_FULL, acc[r][0], o);
            acc[r][1] += __shfl_xor_sync(RR_FULL, acc[r][1], o);
          }
        }
      }
      /* finish the rows leaf-most first; each finished row is folded into the ones above it */
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (r < T) {
          const int dg = mk - r; /* diagonal position of row k - r */
          const float d1 = __shfl_sync(RR_FULL, acc[r][0], dg), d2 = __shfl_sync(RR_FULL, acc[r][1], dg) + m.kdtd[k - r];
          const float l1 = acc[r][0] * RR_RCP(d1), l2 = acc[r][1] * RR_RCP(d2);
#pragma unroll
          for (int q = r + 1; q < 4; q++) {
            if (q < T) {
              const float u1 = __shfl_sync(RR_FULL, acc[r][0], mk - q), u2 = __shfl_sync(RR_FULL, acc[r][1], mk - q);
              acc[q][0] -= u1 * l1; acc[q][1] -= u2 * l2;
            }
          }
          if (g == 0 && s0 <= dg) { /* the lane of the diagonal stores D (inverted at the end of the sweep), the others L */
            LD[adr_r[r] + s0] = s0 == dg ? d1 : l1;
            L2[adr_r[r] + s0] = s0 == dg ? d2 : l2;
          }
        }
      }
      __syncwarp();
      k -= T;
      continue;
    }
    /* rows of 33 .. 64 entries: two registers per lane and matrix, one row per step */
    const float damp = dt * RF(dof_damping, k);
    for (int jj = c.lane; jj < nd; jj += 32) {
      const int pj = RI(dof_pack, k + 1 + jj), rj = pj & 0xffff, dj = (pj >> 16) & 255;
      stA[jj] = make_float4(LD[rj + mk] * LD[rj + dj], L2[rj + mk] * L2[rj + dj], __int_as_float(rj), 0.f);
    }
    {
      const int s0 = c.lane, s1 = c.lane + 32;
      const bool on1 = s1 <= mk;
      const int so1 = on1 ? s1 : 0;
      float a0 = L2[adr + s0], a1 = on1 ? L2[adr + so1] : 0.f, b0 = a0, b1 = a1;
      if (s1 == mk) b1 += damp;
      __syncwarp();
#pragma unroll 2
      for (int jj = 0; jj < nd; jj++) {
        const float4 wr = stA[jj];
        const int rj = __float_as_int(wr.z);
        a0 -= wr.x * LD[rj + s0];
        b0 -= wr.y * L2[rj + s0];
        a1 -= on1 ? wr.x * LD[rj + so1] : 0.f;
        b1 -= on1 ? wr.y * L2[rj + so1] : 0.f;
      }
      const float dk = __shfl_sync(RR_FULL, a1, mk & 31), dk2 = __shfl_sync(RR_FULL, b1, mk & 31);
      const float inv = RR_RCP(dk), inv2 = RR_RCP(dk2);
      LD[adr + s0] = a0 * inv; L2[adr + s0] = b0 * inv2;
      if (s1 < mk) { LD[adr + s1] = a1 * inv; L2[adr + s1] = b1 * inv2; }
      if (c.lane == 0) { LD[adr + mk] = dk; L2[adr + mk] = dk2; }
    }
    __syncwarp();
    k--;
  }
  /* diagonal slots: D -> 1 / D (what the solves multiply by) */
  RR_FOR_S {
    const int i = c.lane + 32 * s;
    float d1 = 0.f, d2 = 0.f;
    if (i < m.nv) {
      const int e = c.radr[s] + c.dep[s];
      d1 = RR_RCP(LD[e]); d2 = RR_RCP(L2[e]);
      LD[e] = d1; L2[e] = d2;
    }
    c.dinv[s] = d1; c.dinv2[s] = d2;
  }
  __syncwarp();
}

/* Right-looking (scatter) variant of factor2, selected with -DRR_FACTOR_RL=1: the same two factorisations, but instead of
 * gathering, per row, the contributions of all its descendants (a reduction behind a dependent stage -> row load chain),
 * each finished row j immediately applies its rank-1 update to the rows of its ancestors:
 *   w = row_j (= L(j, .) D_j),  l = w / D_j,   M(a_q, a_p) -= w_q l_p   for the ancestors a_p <= a_q of j.
 * Lane p owns column position p; the rows a_q are visited in batches of 4 with all loads of a batch issued before its stores
 * (distinct rows: no hazard), so there is no reduction, no shuffle and no dependent load chain -- only the pivot reciprocal
 * is serial per row. */
template <int NS>
RR_DEV void factor2_rl(Ctx<NS> &c, float dt) {
  const RRModelDev &m = c.m;
  float *A1 = c.LD, *A2 = c.M;
  __syncwarp();
  for (int e = c.lane; e < m.nM; e += 32) A1[e] = A2[e];
  __syncwarp();
  RR_FOR_S {
    const int i = c.lane + 32 * s;
    if (i < m.nv) A2[c.radr[s] + c.dep[s]] += dt * RF(dof_damping, i);
  }
  __syncwarp();
  const RR_SADDR_T s1 = RR_SADDR(A1 + c.lane), s2 = RR_SADDR(A2 + c.lane);
#pragma unroll 1
  for (int j = m.nv - 1; j > 0; j--) {
    const int adr4 = c.m.krow4[j], d = c.m.kdep4[j] >> 2;
    if (d == 0) continue;
    const int adr = adr4 >> 2;
    const float i1 = RR_RCP(A1[adr + d]), i2 = RR_RCP(A2[adr + d]);
    const int p0 = c.lane, p1 = c.lane + 32;
    const bool on0 = p0 < d, on1 = p1 < d;
    const float l1a = on0 ? A1[adr + p0] * i1 : 0.f, l2a = on0 ? A2[adr + p0] * i2 : 0.f;
    const float l1b = on1 ? A1[adr + p1] * i1 : 0.f, l2b = on1 ? A2[adr + p1] * i2 : 0.f;
#pragma unroll 1
    for (int q0 = 0; q0 < d; q0 += 4) {
      float t1[4], t2[4], u1[4], u2[4], w1[4], w2[4];
      int rq4[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int q = q0 + k < d ? q0 + k : d - 1;
        rq4[k] = RR_META_ANCRADR(RI(M_meta, adr + q)) << 2;
        w1[k] = A1[adr + q]; w2[k] = A2[adr + q];
        t1[k] = RR_SLOAD(s1 + rq4[k]); t2[k] = RR_SLOAD(s2 + rq4[k]);
        if (d > 32) { u1[k] = RR_SLOAD(s1 + rq4[k] + 128); u2[k] = RR_SLOAD(s2 + rq4[k] + 128); }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int q = q0 + k;
        if (q < d && p0 <= q) { A1[(rq4[k] >> 2) + p0] = t1[k] - w1[k] * l1a; A2[(rq4[k] >> 2) + p0] = t2[k] - w2[k] * l2a; }
        if (d > 32 && q < d && p1 <= q) { A1[(rq4[k] >> 2) + p1] = u1[k] - w1[k] * l1b; A2[(rq4[k] >> 2) + p1] = u2[k] - w2[k] * l2b; }
      }
    }
    __syncwarp();
    if (on0) { A1[adr + p0] = l1a; A2[adr + p0] = l2a; }
    if (on1) { A1[adr + p1] = l1b; A2[adr + p1] = l2b; }
    __syncwarp();
  }
  RR_FOR_S {
    const int i = c.lane + 32 * s;
    float d1 = 0.f, d2 = 0.f;
    if (i < m.nv) {
      const int e = c.radr[s] + c.dep[s];
      d1 = RR_RCP(A1[e]); d2 = RR_RCP(A2[e]);
      A1[e] = d1; A2[e] = d2;
    }
    c.dinv[s] = d1; c.dinv2[s] = d2;
  }
  __syncwarp();
}

/* One dof (column) per step in index order, in groups of 8 columns; per-slot code so that the broadcast register is a
 * compile-time choice, no branch in the body and all metadata / coefficient loads independent of x: the only dependent chain
 * is shuffle -> FMA (-> next shuffle).  Per column and slot the work is: test one bit of a static mask (is this column a
 * descendant / an ancestor of my dof -- host tables dof_descmask / dof_ancmask, shifted once per group), one add for the
 * coefficient address (per-lane shared-memory byte address + the column's byte offset from the constant bank), the load
 * and the FMA.  Columns beyond nv have empty masks and do nothing (the last slot runs whole groups of 8). */
template <int NS>
RR_DEV void solve_ld(Ctx<NS> &c, float (&x)[NS], const float *LDm, const float (&dinv)[NS]) {
  const int nv = c.m.nv;
  const int nb32 = (nv + 31) >> 5;
  RR_SADDR_T lcol[NS], lrow[NS]; /* byte address of L(., lane dof) in row 0; of the row of the lane dof */
  RR_FOR_S { lcol[s] = RR_SADDR(LDm + c.dep[s]); lrow[s] = RR_SADDR(LDm + c.radr[s]); }
  /* backward: x <- L^-T x, leaves to root; column i updates its ancestors = the dofs that have i as a descendant */
#pragma unroll
  for (int si = NS - 1; si >= 0; si--) {
    if (32 * si >= nv) continue;
    unsigned dm[NS];
#pragma unroll
    for (int s = 0; s <= si; s++) {
      const int j = c.lane + 32 * s;
      dm[s] = j < nv ? (unsigned)RI(dof_descmask, j * nb32 + si) : 0u;
    }
    const int ng = si == NS - 1 ? (((nv - 32 * si) < 32 ? (nv - 32 * si) : 32) + 7) >> 3 : 4;
#pragma unroll 1
    for (int g = ng - 1; g >= 0; g--) {
      unsigned mg[NS];
#pragma unroll
      for (int s = 0; s <= si; s++) mg[s] = dm[s] >> (8 * g);
#pragma unroll
      for (int k = 7; k >= 0; k--) {
        const int src = 8 * g + k;
        const int adr4 = c.m.krow4[32 * si + src];
        const float xi = __shfl_sync(RR_FULL, x[si], src);
#pragma unroll
        for (int s = 0; s <= si; s++) {
          RR_SOLVE_UPDATE(x[s], mg[s], 1u << k, lcol[s] + adr4, xi);
        }
      }
    }
  }
  RR_FOR_S x[s] *= dinv[s];
  /* forward: x <- L^-1 x, root to leaves; column j updates its descendants = the dofs that have j as an ancestor */
#pragma unroll
  for (int sj = 0; sj < NS; sj++) {
    if (32 * sj >= nv) continue;
    unsigned am[NS];
#pragma unroll
    for (int s = sj; s < NS; s++) {
      const int i = c.lane + 32 * s;
      am[s] = i < nv ? (unsigned)RI(dof_ancmask, i * nb32 + sj) : 0u;
    }
    const int ng = sj == NS - 1 ? (((nv - 32 * sj) < 32 ? (nv - 32 * sj) : 32) + 7) >> 3 : 4;
#pragma unroll 1
    for (int g = 0; g < ng; g++) {
      unsigned mg[NS];
#pragma unroll
      for (int s = sj; s < NS; s++) mg[s] = am[s] >> (8 * g);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int src = 8 * g + k;
        const int dp4 = c.m.kdep4[32 * sj + src];
        const float xj = __shfl_sync(RR_FULL, x[sj], src);
#pragma unroll
        for (int s = sj; s < NS; s++) {
          RR_SOLVE_UPDATE(x[s], mg[s], 1u << k, lrow[s] + dp4, xj);
        }
      }
    }
  }
}

/* y = M v (symmetric tree-sparse product); v is staged through vbuf */
template <int NS>
RR_DEV void mul_m(Ctx<NS> &c, float (&y)[NS], const float (&v)[NS]) {
  const RRModelDev &m = c.m;
  __syncwarp();
  vstore<NS>(c, v, c.vbuf);
  __syncwarp();
  RR_FOR_S {
    int i = c.lane + 32 * s;
    float acc = 0.f;
    if (i < m.nv) {
      int adr = c.radr[s];
#pragma unroll 8
      for (int t = 0; t <= c.dep[s]; t++) acc += c.M[adr + t] * c.vbuf[RR_META_COL(RI(M_meta, adr + t))];
#pragma unroll 8
      for (int k = i + 1; k <= i + c.nd[s]; k++) acc += c.M[RI(dof_rowadr, k) + c.dep[s]] * c.vbuf[k];
    }
    y[s] = acc;
  }
  __syncwarp();
}

/* ------------------------------------------------------------------------------------------ velocity + rne (B.6) */
/* The forward scans (cvel, cacc) are prefix SUMS over the chain of a body (everything is expressed about the tree COM, so no
 * transforms are involved): like the kinematics they run by pointer doubling over body_anc instead of level by level.
 *   A_b = sum over the joints of b of cdof qvel;  S = inclusive prefix of A;  the velocity entering b is S[eparent(b)];
 *   the per-joint sequence of mjx com_vel (cdof_dot = cvel x cdof with the velocity accumulated so far) then runs per body
 *   in parallel and leaves cvel_b and W_b = sum of cdof_dot qvel;  cacc_b = -gravity + inclusive prefix of W. */
template <int NS>
RR_DEV void tree_prefix6(Ctx<NS> &c, float *&in, float *&out) {
  const RRModelDev &m = c.m;
  const int nb = m.nbody;
#pragma unroll 1
  for (int rd = 0; rd < m.nround; rd++) {
    for (int b = c.lane; b < nb; b += 32) {
      const int an = RR_BODY_ANC(rd * nb + b);
      const float2 *ia = reinterpret_cast<const float2 *>(in + 6 * an), *ib = reinterpret_cast<const float2 *>(in + 6 * b);
      float2 *ob = reinterpret_cast<float2 *>(out + 6 * b);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float2 u = ia[k], v = ib[k];
        ob[k] = make_float2(u.x + v.x, u.y + v.y);
      }
    }
    __syncwarp();
    float *t = in; in = out; out = t;
  }
}

template <int NS>
RR_DEV void com_vel_and_rne(Ctx<NS> &c, float (&qfrc_bias)[NS]) {
  const RRModelDev &m = c.m;
  const int nb = m.nbody;
  const bool odd = m.nround & 1;
  __syncwarp();
  /* A_b -> the buffer from which nround swaps end in LD (so that cvel can be written while S is still being read) */
  float *in = odd ? c.cvel : c.LD, *out = odd ? c.LD : c.cvel;
  for (int b = c.lane; b < nb; b += 32) {
    float av[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int jadr = RI(body_jntadr, b), jnum = RI(body_jntnum, b);
    for (int j = jadr; j < jadr + jnum; j++) {
      const int d0 = RI(jnt_dofadr, j), nd = RI(jnt_type, j) == RR_JNT_FREE ? 6 : 1;
      for (int d = 0; d < nd; d++) {
        const float qv = c.qvel[d0 + d];
#pragma unroll
        for (int k = 0; k < 6; k++) av[k] += c.cdof[6 * (d0 + d) + k] * qv;
      }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) in[6 * b + k] = av[k];
  }
  __syncwarp();
  tree_prefix6<NS>(c, in, out); /* `in` = LD now holds S */
  float *win = c.cacc, *wout = c.LD; /* W_b starts in cacc (S in LD is still being read by the other lanes) */
  const float *S = in;
  for (int b = c.lane; b < nb; b += 32) {
    const int p = RI(body_eparent, b);
    float cv[6], w[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 6; k++) cv[k] = S[6 * p + k];
    const int jadr = RI(body_jntadr, b), jnum = RI(body_jntnum, b);
    for (int j = jadr; j < jadr + jnum; j++) {
      const int d0 = RI(jnt_dofadr, j);
      if (RI(jnt_type, j) == RR_JNT_FREE) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const float qv = c.qvel[d0 + d];
#pragma unroll
          for (int k = 0; k < 6; k++) cv[k] += c.cdof[6 * (d0 + d) + k] * qv;
        }
        float cdd[3][6];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          float cd[6];
#pragma unroll
          for (int k = 0; k < 6; k++) cd[k] = c.cdof[6 * (d0 + 3 + d) + k];
          motion_cross(cdd[d], cv, cd);
        }
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const float qv = c.qvel[d0 + 3 + d];
#pragma unroll
          for (int k = 0; k < 6; k++) { cv[k] += c.cdof[6 * (d0 + 3 + d) + k] * qv; w[k] += cdd[d][k] * qv; }
        }
      } else {
        float cd[6], cdd[6];
        const float qv = c.qvel[d0];
#pragma unroll
        for (int k = 0; k < 6; k++) cd[k] = c.cdof[6 * d0 + k];
        motion_cross(cdd, cv, cd);
#pragma unroll
        for (int k = 0; k < 6; k++) { cv[k] += cd[k] * qv; w[k] += cdd[k] * qv; }
      }
    }
    if (b == 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) cv[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) { c.cvel[6 * b + k] = cv[k]; win[6 * b + k] = w[k]; }
  }
  __syncwarp();
  tree_prefix6<NS>(c, win, wout); /* `win` (cacc or LD) now holds the inclusive prefix of W */
  for (int i = c.lane; i < 6 * nb; i += 32) {
    const int k = i % 6;
    c.cacc[i] = win[i] - (k >= 3 ? m.gravity[k - 3] : 0.f);
  }
  __syncwarp();
  /* local body forces */
  for (int b = c.lane; b < m.nbody; b += 32) {
    float ci[10], cv[6], ca[6], f1[6], f2[6], f3[6];
#pragma unroll
    for (int k = 0; k < 10; k++) ci[k] = c.cinert[10 * b + k];
#pragma unroll
    for (int k = 0; k < 6; k++) { cv[k] = c.cvel[6 * b + k]; ca[k] = c.cacc[6 * b + k]; }
    inert_mul(f1, ci, ca);
    inert_mul(f2, ci, cv);
    motion_cross_force(f3, cv, f2);
#pragma unroll
    for (int k = 0; k < 6; k++) c.cfrc[6 * b + k] = f1[k] + f3[k]; /* cfrc aliases cacc: body-local, in place */
  }
  __syncwarp();
#pragma unroll 4
  for (int b = m.nbody - 1; b > 0; b--) { /* lane q owns component q: no rendezvous needed inside the loop */
    const int p = m.kpar[b];
    if (c.lane < 6 && p > 0) c.cfrc[6 * p + c.lane] += c.cfrc[6 * b + c.lane];
  }
  __syncwarp();
  RR_FOR_S {
    int i = c.lane + 32 * s;
    float v = 0.f;
    if (i < m.nv) {
      int b = RI(dof_bodyid, i);
#pragma unroll
      for (int k = 0; k < 6; k++) v += c.cdof[6 * i + k] * c.cfrc[6 * b + k];
    }
    qfrc_bias[s] = v;
  }
}

/* y = M v from the bodies' own inertias instead of the assembled matrix: M = sum_b J_b' I_b J_b (+ armature), so
 *   A_b = prefix sum over the chain of b of cdof v   (tree_prefix6, as com_vel),   F_b = I_b A_b  (cinert, inert_mul),
 *   y_d = cdof_d . (sum of F over the subtree of body(d)) + armature_d v_d          (leaf-to-root accumulation, as rne).
 * Same bilinear form as crb_and_mass_matrix builds, O(nbody) work on all lanes instead of a row + column walk per dof in
 * which the root dofs alone visit 72 descendants (mul_m: 1.5 k instructions per call, 6 busy lanes for most of them).
 * Needs cinert and cdof of the current pose and uses cvel / cacc / LD as scratch: only valid between smooth_forces and the
 * factorisation (the M qacc_warmstart product of the solver's warm start). */
template <int NS>
RR_DEV void mul_m_tree(Ctx<NS> &c, float (&y)[NS], const float (&v)[NS]) {
  const RRModelDev &m = c.m;
  const int nb = m.nbody;
  const bool odd = m.nround & 1;
  __syncwarp();
  vstore<NS>(c, v, c.vbuf);
  __syncwarp();
  float *in = odd ? c.cvel : c.LD, *out = odd ? c.LD : c.cvel; /* nround swaps end in LD */
  for (int b = c.lane; b < nb; b += 32) {
    float av[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int jadr = RI(body_jntadr, b), jnum = RI(body_jntnum, b);
    for (int j = jadr; j < jadr + jnum; j++) {
      const int d0 = RI(jnt_dofadr, j), nd = RI(jnt_type, j) == RR_JNT_FREE ? 6 : 1;
      for (int d = 0; d < nd; d++) {
        const float qv = c.vbuf[d0 + d];
#pragma unroll
        for (int k = 0; k < 6; k++) av[k] += c.cdof[6 * (d0 + d) + k] * qv;
      }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) in[6 * b + k] = av[k];
  }
  __syncwarp();
  tree_prefix6<NS>(c, in, out); /* `in` = LD now holds A_b */
  for (int b = c.lane; b < nb; b += 32) {
    float ci[10], a6[6], f[6];
#pragma unroll
    for (int k = 0; k < 10; k++) ci[k] = c.cinert[10 * b + k];
#pragma unroll
    for (int k = 0; k < 6; k++) a6[k] = in[6 * b + k];
    inert_mul(f, ci, a6);
#pragma unroll
    for (int k = 0; k < 6; k++) c.cfrc[6 * b + k] = f[k];
  }
  __syncwarp();
#pragma unroll 4
  for (int b = nb - 1; b > 0; b--) { /* lane q owns component q */
    const int p = m.kpar[b];
    if (c.lane < 6 && p > 0) c.cfrc[6 * p + c.lane] += c.cfrc[6 * b + c.lane];
  }
  __syncwarp();
  RR_FOR_S {
    const int i = c.lane + 32 * s;
    float acc = 0.f;
    if (i < m.nv) {
      const int b = RI(dof_bodyid, i);
      acc = RF(dof_armature, i) * v[s];
#pragma unroll
      for (int k = 0; k < 6; k++) acc += c.cdof[6 * i + k] * c.cfrc[6 * b + k];
    }
    y[s] = acc;
  }
  __syncwarp();
}

/* passive + actuation -> qfrc_smooth (qacc_smooth = M^-1 qfrc_smooth is solved by the caller) */
template <int NS>
RR_DEV void smooth_forces(Ctx<NS> &c, const float (&qfrc_bias)[NS]) {
  const RRModelDev &m = c.m;
  __syncwarp();
  /* passive springs into vbuf (per joint), dampers added per dof below */
  for (int i = c.lane; i < m.nv; i += 32) { c.vbuf[i] = 0.f; c.qfrc_act[i] = 0.f; }
  __syncwarp();
  for (int j = c.lane; j < m.njnt; j += 32) {
    int qa = RI(jnt_qposadr, j), da = RI(jnt_dofadr, j);
    float k = RF(jnt_stiffness, j);
    if (RI(jnt_type, j) == RR_JNT_FREE) {
      /* free-joint spring: translational part only matters when stiffness != 0 (never for <freejoint>) */
#pragma unroll
      for (int d = 0; d < 3; d++) c.vbuf[da + d] = -k * (c.qpos[qa + d] - RF(qpos_spring, qa + d));
    } else {
      c.vbuf[da] = -k * (c.qpos[qa] - RF(qpos_spring, qa));
    }
  }
  /* actuation (fwd_actuation): filter activation, affine gain / bias, joint transmission */
  for (int u = c.lane; u < m.nu; u += 32) {
    float ctrl = c.ctrl[u];
    if (RI(act_ctrllimited, u)) ctrl = clampf(ctrl, RF(act_ctrlrange, 2 * u), RF(act_ctrlrange, 2 * u + 1));
    int da = RI(act_dofadr, u), qa = RI(act_qposadr, u);
    float gear = RF(act_gear, u);
    float len = gear * c.qpos[qa], vel = gear * c.qvel[da];
    float ctrl_act = ctrl;
    if (RI(act_dyntype, u) == 2) {
      int aa = RI(act_actadr, u);
      float tau = fmaxf(RF(act_dynprm, u), RR_MINVAL);
      c.actdot[aa] = (ctrl - c.act[aa]) * RR_RCP(tau);
      ctrl_act = c.act[aa];
    }
    float gain = RF(act_gainprm, 3 * u);
    if (RI(act_gaintype, u) == 1) gain += RF(act_gainprm, 3 * u + 1) * len + RF(act_gainprm, 3 * u + 2) * vel;
    float bias = 0.f;
    if (RI(act_biastype, u) == 1)
      bias = RF(act_biasprm, 3 * u) + RF(act_biasprm, 3 * u + 1) * len + RF(act_biasprm, 3 * u + 2) * vel;
    float force = gain * ctrl_act + bias;
    if (RI(act_forcelimited, u)) force = clampf(force, RF(act_forcerange, 2 * u), RF(act_forcerange, 2 * u + 1));
    c.qfrc_act[da] = gear * force;
  }
  __syncwarp();
  float passive[NS];
  RR_FOR_S {
    int i = c.lane + 32 * s;
    float pv = 0.f, av = 0.f;
    if (i < m.nv) { pv = c.vbuf[i] - RF(dof_damping, i) * c.qvel[i]; av = c.qfrc_act[i]; }
    passive[s] = pv;
    c.qfrc_smooth[s] = pv - qfrc_bias[s] + av;
    c.qacc_smooth[s] = c.qfrc_smooth[s];
  }
  dbg_vec<NS>(c, RR_DBG_QFRC_BIAS, qfrc_bias);
  dbg_vec<NS>(c, RR_DBG_QFRC_PASSIVE, passive);
  dbg_copy<NS>(c, RR_DBG_QFRC_ACTUATOR, c.qfrc_act, m.nv);
  dbg_vec<NS>(c, RR_DBG_QFRC_SMOOTH, c.qfrc_smooth);
}

/* ------------------------------------------------------------------------------------------ collision (B.4) */
RR_DEV void make_frame(float *frame, const float *n) {
  float a[3] = {n[0], n[1], n[2]};
  normalize3(a);
  float b[3] = {0.f, 0.f, 0.f};
  if (-0.5f < a[1] && a[1] < 0.5f) b[1] = 1.f; else b[2] = 1.f;
  float ab = dot3(a, b);
#pragma unroll
  for (int k = 0; k < 3; k++) b[k] -= a[k] * ab;
  normalize3(b);
#pragma unroll
  for (int k = 0; k < 3; k++) { frame[k] = a[k]; frame[3 + k] = b[k]; }
  cross3(frame + 6, a, b);
}

template <int NS>
RR_DEV void collision(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  for (int p = c.lane; p < m.npair; p += 32) {
    int b = RI(pair_body, p), ca = RI(pair_conadr, p), fn = RI(pair_fn, p);
    const float *cm = c.com + 3 * RI(body_rootslot, b); /* contact points are kept relative to the tree COM (what cdof refers to) */
    float n[3], pp[3], gl[3], gq[4], xq[4], gp[3], r[3], size[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      n[k] = RF(pair_plane_n, 3 * p + k); pp[k] = RF(pair_plane_p, 3 * p + k);
      gl[k] = RF(pair_gpos, 3 * p + k); size[k] = RF(pair_size, 3 * p + k);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) { xq[k] = c.xquat[4 * b + k]; gq[k] = RF(pair_gquat, 4 * p + k); }
    rotq(r, gl, xq);
#pragma unroll
    for (int k = 0; k < 3; k++) gp[k] = c.xpos[3 * b + k] + r[k];
    if (fn == RR_PAIR_PLANE_SPHERE) {
      float d[3] = {gp[0] - pp[0], gp[1] - pp[1], gp[2] - pp[2]};
      float dist = dot3(d, n) - size[0];
      c.con_dist[ca] = dist;
#pragma unroll
      for (int k = 0; k < 3; k++) c.cab[12 * ca + 9 + k] = gp[k] - n[k] * (size[0] + 0.5f * dist) - cm[k];
      make_frame(c.cab + 12 * ca, n);
    } else {
      float q[4], gm[9];
      quat_mul(q, xq, gq);
      quat_to_mat(gm, q);
      if (fn == RR_PAIR_PLANE_CAPSULE) {
        float axis[3] = {gm[2], gm[5], gm[8]};
        float na = dot3(n, axis), bv[3], frame[9];
#pragma unroll
        for (int k = 0; k < 3; k++) bv[k] = axis[k] - n[k] * na;
        float bn = normalize3(bv);
        if (bn < 0.5f) {
          bv[0] = 0.f;
          if (-0.5f < n[1] && n[1] < 0.5f) { bv[1] = 1.f; bv[2] = 0.f; } else { bv[1] = 0.f; bv[2] = 1.f; }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) { frame[k] = n[k]; frame[3 + k] = bv[k]; }
        cross3(frame + 6, n, bv);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          float sg = e == 0 ? 1.f : -1.f, cp[3], d[3];
#pragma unroll
          for (int k = 0; k < 3; k++) { cp[k] = gp[k] + sg * (axis[k] * size[1]); d[k] = cp[k] - pp[k]; }
          float dist = dot3(d, n) - size[0];
          c.con_dist[ca + e] = dist;
#pragma unroll
          for (int k = 0; k < 3; k++) c.cab[12 * (ca + e) + 9 + k] = cp[k] - n[k] * (size[0] + 0.5f * dist) - cm[k];
#pragma unroll
          for (int k = 0; k < 9; k++) c.cab[12 * (ca + e) + k] = frame[k];
        }
      } else { /* plane - ellipsoid */
        float nl[3], sv[3], lp[3], wp[3];
#pragma unroll
        for (int k = 0; k < 3; k++) nl[k] = gm[k] * n[0] + gm[3 + k] * n[1] + gm[6 + k] * n[2];
#pragma unroll
        for (int k = 0; k < 3; k++) sv[k] = nl[k] * size[k];
        float nrm = sqrtf(dot3(sv, sv));
#pragma unroll
        for (int k = 0; k < 3; k++) lp[k] = -(sv[k] / nrm) * size[k];
#pragma unroll
        for (int k = 0; k < 3; k++) wp[k] = gp[k] + gm[3 * k] * lp[0] + gm[3 * k + 1] * lp[1] + gm[3 * k + 2] * lp[2];
        float d[3] = {wp[0] - pp[0], wp[1] - pp[1], wp[2] - pp[2]};
        float dist = dot3(d, n);
        c.con_dist[ca] = dist;
#pragma unroll
        for (int k = 0; k < 3; k++) c.cab[12 * ca + 9 + k] = wp[k] - n[k] * dist * 0.5f - cm[k];
        make_frame(c.cab + 12 * ca, n);
      }
    }
  }
  __syncwarp();
  dbg_copy<NS>(c, RR_DBG_CON_DIST, c.con_dist, m.ncon);
  if (RR_WITH_DEBUG && c.dbg) {
    float *dP = c.dbg + dbg_offset(m, RR_DBG_CON_POS), *dF = c.dbg + dbg_offset(m, RR_DBG_CON_FRAME);
    for (int i = c.lane; i < 3 * m.ncon; i += 32) {
      int cc = i / 3, k = i % 3;
      dP[i] = c.cab[12 * cc + 9 + k] + c.com[3 * RI(body_rootslot, RI(pair_body, RI(con_pair, cc))) + k];
    }
    for (int i = c.lane; i < 9 * m.ncon; i += 32) dF[i] = c.cab[12 * (i / 9) + i % 9];
  }
}

/* ------------------------------------------------------------------------------------------ constraint rows (B.5) */
/* Contact Jacobians are never materialised.  A contact row in frame direction r is
 *   J_r[d] = fr_r . (cdof_lin[d] + cdof_ang[d] x off) = [off x fr_r ; fr_r] . cdof[d]        (off = contact point - tree COM)
 * Each contact keeps its frame and off (cab: 12 floats), and
 *   J v   : V_b = sum over the chain of body b of cdof[d] v[d] (spatial velocity induced by v), u = V_lin + V_ang x off,
 *           (J v)_r = fr_r . u
 *   J' f  : per contact the world force w = sum_r g_r fr_r and its moment off x w about the tree COM (six floats, cscr),
 *           F_b = their sum over the contacts on b,  qfc[d] = cdof[d] . F_b for d in chain(b)
 * with one V_b / F_b per contact BODY (10 for the rodent's 34 contacts): work no longer scales with chain x contacts.
 * cscr holds six floats per ACTIVE contact (compact index), in shared memory up to capA contacts. */

/* rows: out[r] = J_r . v for the compact active rows; v staged in vbuf by the caller (already synced). */
template <int NS>
RR_DEV void mul_j(Ctx<NS> &c, float *out) {
  const RRModelDev &m = c.m;
  for (int r = c.lane; r < c.nla; r += 32) {
    int id = c.row_id[r];
    float sg = (id & RR_SIGN_BIT) ? -1.f : 1.f;
    out[r] = sg * c.vbuf[(id >> 16) & 255];
  }
  /* V_b for every contact body: item = (body slot, component) */
  for (int it = c.lane; it < 6 * m.ncb; it += 32) {
    int kb = it / 6, q = it - 6 * kb;
    int ld = RI(cb_lastdof, kb);
    int len = RI(dof_depth, ld) + 1, adr = RI(dof_rowadr, ld);
    float acc = 0.f;
#pragma unroll 4
    for (int t = 0; t < len; t++) {
      int d = RR_META_COL(RI(M_meta, adr + t));
      acc += c.cdof[6 * d + q] * c.vbuf[d];
    }
    c.cbv[it] = acc;
  }
  __syncwarp();
  for (int k = c.lane; k < c.nca; k += 32) {
    const int ck = c.cact[k], cc = ck & 255; /* cact: contact | pair << 8 | contact body << 20 (packed by make_constraint) */
    const float *ab = c.cab + 12 * cc, *V = c.cbv + 6 * (ck >> 20);
    const float off[3] = {ab[9], ab[10], ab[11]}, va[3] = {V[0], V[1], V[2]};
    float u[3];
    cross3(u, va, off);
#pragma unroll
    for (int q = 0; q < 3; q++) u[q] += V[3 + q];
#pragma unroll
    for (int r3 = 0; r3 < 3; r3++) c.cscr[6 * k + r3] = ab[3 * r3] * u[0] + ab[3 * r3 + 1] * u[1] + ab[3 * r3 + 2] * u[2];
  }
  __syncwarp();
  for (int r = c.lane; r < 4 * c.nca; r += 32) {
    int k = r >> 2, q = r & 3;
    float mu = RF(pair_mu, (c.cact[k] >> 8) & 4095);
    float f = (q & 1) ? -mu : mu;
    out[c.nla + r] = c.cscr[6 * k] + c.cscr[6 * k + 1 + (q >> 1)] * f;
  }
  __syncwarp();
}

/* qfc = J' f for the compact active rows; f in `frc` (rows) */
template <int NS>
RR_DEV void mul_jt(Ctx<NS> &c, const float *frc, float (&qfc)[NS]) {
  const RRModelDev &m = c.m;
  __syncwarp();
  for (int i = c.lane; i < m.nv; i += 32) c.vbuf[i] = 0.f;
  __syncwarp();
  for (int r = c.lane; r < c.nla; r += 32) {
    int id = c.row_id[r];
    float sg = (id & RR_SIGN_BIT) ? -1.f : 1.f;
    c.vbuf[(id >> 16) & 255] = sg * frc[r];
  }
  for (int k = c.lane; k < c.nca; k += 32) {
    const float *f = frc + c.nla + 4 * k;
    const int ck = c.cact[k], cc = ck & 255;
    const float mu = RF(pair_mu, (ck >> 8) & 4095);
    const float g0 = f[0] + f[1] + f[2] + f[3], g1 = mu * f[0] - mu * f[1], g2 = mu * f[2] - mu * f[3];
    const float *ab = c.cab + 12 * cc;
    float w[3], t[3];
#pragma unroll
    for (int q = 0; q < 3; q++) w[q] = g0 * ab[q] + g1 * ab[3 + q] + g2 * ab[6 + q];
    const float off[3] = {ab[9], ab[10], ab[11]};
    cross3(t, off, w);
#pragma unroll
    for (int q = 0; q < 3; q++) { c.cscr[6 * k + q] = t[q]; c.cscr[6 * k + 3 + q] = w[q]; }
  }
  __syncwarp();
  /* F_b: item = (body slot, component), summed over that body's (static) contact list; ckidx maps a contact to its
   * compact active index (-1 = inactive) */
  for (int it = c.lane; it < 6 * m.ncb; it += 32) {
    int kb = it / 6, q = it - 6 * kb;
    float acc = 0.f;
    const int beg = RI(cb_conadr, kb), end = RI(cb_conadr, kb + 1);
#pragma unroll 2
    for (int e = beg; e < end; e++) {
      int cc = RI(cb_conlist, e);
      int k = c.ckidx[cc];
      if (k >= 0) acc += c.cscr[6 * k + q];
    }
    c.cbv[it] = acc;
  }
  __syncwarp();
  RR_FOR_S {
    int i = c.lane + 32 * s;
    float acc = 0.f;
    if (i < m.nv) {
      acc = c.vbuf[i];
      const float *cd = c.cdof + 6 * i;
      /* contact bodies whose chain contains dof i: static 64-bit mask in two words */
#pragma unroll
      for (int w = 0; w < 2; w++) {
        unsigned mask = (unsigned)RI(dof_cbmask, 2 * i + w);
        while (mask) {
          const int kb = 32 * w + __ffs(mask) - 1;
          mask &= mask - 1;
          const float *F = c.cbv + 6 * kb;
          acc += cd[0] * F[0] + cd[1] * F[1] + cd[2] * F[2] + cd[3] * F[3] + cd[4] * F[4] + cd[5] * F[5];
        }
      }
    }
    qfc[s] = acc;
  }
  __syncwarp();
}

template <int NS>
RR_DEV void make_constraint(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  const unsigned lt = (1u << c.lane) - 1u;
  /* pass 1: count the active limit rows and contacts, then pick the row storage (shared memory up to capR rows) */
  int nla = 0, nca = 0;
  for (int base = 0; base < m.nlimit; base += 32) {
    int l = base + c.lane;
    bool active = false;
    if (l < m.nlimit) {
      float q = c.qpos[RI(limit_qposadr, l)];
      float pos = fminf(q - RF(limit_range, 2 * l), RF(limit_range, 2 * l + 1) - q) - RF(limit_margin, l);
      active = pos < 0.f;
    }
    nla += __popc(__ballot_sync(RR_FULL, active));
  }
  for (int base = 0; base < m.ncon; base += 32) {
    int cc = base + c.lane;
    bool active = false;
    if (cc < m.ncon) active = (c.con_dist[cc] - RF(pair_margin, RI(con_pair, cc))) < 0.f;
    unsigned mask = __ballot_sync(RR_FULL, active);
    if (cc < m.ncon) {
      int k = nca + __popc(mask & lt);
      if (active) { const int pp = RI(con_pair, cc); c.cact[k] = cc | (pp << 8) | (RI(pair_cb, pp) << 20); }
      c.ckidx[cc] = active ? k : -1;
    }
    nca += __popc(mask);
  }
  const int nra = nla + 4 * nca;
  c.nla = nla;
  c.nca = nca;
  c.use_rows(nra <= m.sm.capR);
  c.cscr = nca <= m.sm.capA ? c.sm_base + m.sm.cscr : c.grows + 5 * m.nefc + 4; /* six floats per active contact */
  __syncwarp();
  /* pass 2: limit rows */
  int r0 = 0;
  for (int base = 0; base < m.nlimit; base += 32) {
    int l = base + c.lane;
    bool active = false;
    float pos = 0.f, dlo = 0.f, dhi = 0.f;
    if (l < m.nlimit) {
      float q = c.qpos[RI(limit_qposadr, l)];
      dlo = q - RF(limit_range, 2 * l);
      dhi = RF(limit_range, 2 * l + 1) - q;
      pos = fminf(dlo, dhi) - RF(limit_margin, l);
      active = pos < 0.f;
    }
    unsigned mask = __ballot_sync(RR_FULL, active);
    if (active) {
      int r = r0 + __popc(mask & lt);
      float sr[2] = {RF(limit_solref, 2 * l), RF(limit_solref, 2 * l + 1)}, si[5], k, b, imp;
#pragma unroll
      for (int q = 0; q < 5; q++) si[q] = RF(limit_solimp, 5 * l + q);
      kbi(m.timestep, sr, si, pos, k, b, imp);
      float R = fmaxf(RF(limit_invweight, l) * (1.f - imp) * RR_RCP(imp), RR_MINVAL);
      c.row_id[r] = l | (RI(limit_dofadr, l) << 16) | (dlo < dhi ? 0 : RR_SIGN_BIT); /* row | dof << 16 | sign */
      c.row_D[r] = RR_RCP(R);
      c.row_aref[r] = k * imp * pos; /* temp: completed below */
      c.row_Jaref[r] = b;            /* temp */
    }
    r0 += __popc(mask);
  }
  /* active contacts: row parameters (the frame and the point stay as collision() left them) */
  for (int k = c.lane; k < nca; k += 32) {
    int cc = c.cact[k] & 255;
    int p = RI(con_pair, cc);
    float pos = c.con_dist[cc] - RF(pair_margin, p);
    float sr[2] = {RF(pair_solref, 2 * p), RF(pair_solref, 2 * p + 1)}, si[5], kk, b, imp;
#pragma unroll
    for (int q = 0; q < 5; q++) si[q] = RF(pair_solimp, 5 * p + q);
    kbi(m.timestep, sr, si, pos, kk, b, imp);
    float mu = RF(pair_mu, p), t = RF(pair_invweight, p);
    float invw = (t + mu * mu * t) * 2.f * mu * mu / m.impratio;
    float R = fmaxf(invw * (1.f - imp) * RR_RCP(imp), RR_MINVAL);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int r = nla + 4 * k + q;
      c.row_id[r] = m.nlimit + 4 * cc + q;
      c.row_D[r] = RR_RCP(R);
      c.row_aref[r] = kk * imp * pos;
      c.row_Jaref[r] = b;
    }
  }
  __syncwarp();
  /* debug: dense efc_J / efc_D in MJX row order (inactive rows are zero there) */
  if (RR_WITH_DEBUG && c.dbg) {
    float *dJ = c.dbg + dbg_offset(m, RR_DBG_EFC_J), *dD = c.dbg + dbg_offset(m, RR_DBG_EFC_D);
    for (int i = c.lane; i < m.nefc * m.nv; i += 32) dJ[i] = 0.f;
    for (int i = c.lane; i < m.nefc; i += 32) dD[i] = 1.f / RR_MINVAL;
    __syncwarp();
    for (int r = c.lane; r < nla; r += 32) {
      int id = c.row_id[r], l = id & 0xffff;
      dJ[l * m.nv + RI(limit_dofadr, l)] = (id & RR_SIGN_BIT) ? -1.f : 1.f;
      dD[l] = c.row_D[r];
    }
    for (int k = 0; k < nca; k++) {
      int cc = c.cact[k] & 255;
      int p = RI(con_pair, cc);
      int ld = RI(pair_lastdof, p);
      int len = RI(dof_depth, ld) + 1, adr = RI(dof_rowadr, ld);
      const float *ab = c.cab + 12 * cc;
      const float off[3] = {ab[9], ab[10], ab[11]};
      float mu = RF(pair_mu, p);
      for (int t = c.lane; t < len; t += 32) {
        int d = RR_META_COL(RI(M_meta, adr + t));
        const float *cd = c.cdof + 6 * d;
        float j3[3];
        for (int r3 = 0; r3 < 3; r3++) {
          float x[3];
          cross3(x, off, ab + 3 * r3);
          j3[r3] = 0.f;
          for (int q = 0; q < 3; q++) j3[r3] += x[q] * cd[q] + ab[3 * r3 + q] * cd[3 + q];
        }
        for (int q = 0; q < 4; q++) {
          float f = (q & 1) ? -mu : mu;
          dJ[(m.nlimit + 4 * cc + q) * m.nv + d] = j3[0] + j3[1 + (q >> 1)] * f;
        }
      }
      if (c.lane < 4) dD[m.nlimit + 4 * cc + c.lane] = c.row_D[nla + 4 * k + c.lane];
    }
    __syncwarp();
  }
  /* aref = -b (J qvel) - k imp pos */
  for (int i = c.lane; i < m.nv; i += 32) c.vbuf[i] = c.qvel[i];
  __syncwarp();
  mul_j<NS>(c, c.row_jv);
  for (int r = c.lane; r < nra; r += 32) c.row_aref[r] = -c.row_Jaref[r] * c.row_jv[r] - c.row_aref[r];
  __syncwarp();
  if (RR_WITH_DEBUG && c.dbg) {
    float *dA = c.dbg + dbg_offset(m, RR_DBG_EFC_AREF);
    for (int i = c.lane; i < m.nefc; i += 32) dA[i] = 0.f;
    __syncwarp();
    for (int r = c.lane; r < nra; r += 32) dA[c.row_id[r] & 0xffff] = c.row_aref[r];
    __syncwarp();
  }
}

/* ------------------------------------------------------------------------------------------ solver (B.7) */
struct LSPoint { float alpha, cost, d0, d1; };

/* _LSPoint.create at NA step sizes in one pass over the active rows (the three candidates of a line-search
 * iteration share the row loads; their 3 NA partial sums are reduced with interleaved shuffles). */
template <int NS, int NA>
RR_DEV void ls_eval(Ctx<NS> &c, int nra, const float (&alpha)[NA], float g0, float g1, float g2, LSPoint (&out)[NA]) {
  float q[NA][3];
#pragma unroll
  for (int k = 0; k < NA; k++) q[k][0] = q[k][1] = q[k][2] = 0.f;
  for (int r = c.lane; r < nra; r += 32) {
    float ja = c.row_Jaref[r], jv = c.row_jv[r], D = c.row_D[r];
    float a0 = 0.5f * ja * ja * D, a1 = jv * ja * D, a2 = 0.5f * jv * jv * D;
#pragma unroll
    for (int k = 0; k < NA; k++) {
      if (ja + alpha[k] * jv < 0.f) { q[k][0] += a0; q[k][1] += a1; q[k][2] += a2; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < NA; k++) {
#pragma unroll
      for (int d = 0; d < 3; d++) q[k][d] += __shfl_xor_sync(RR_FULL, q[k][d], o);
    }
  }
#pragma unroll
  for (int k = 0; k < NA; k++) {
    float q0 = g0 + q[k][0], q1 = g1 + q[k][1], q2 = g2 + q[k][2], al = alpha[k];
    out[k].alpha = al;
    out[k].cost = al * al * q2 + al * q1 + q0;
    out[k].d0 = 2.f * al * q2 + q1;
    out[k].d1 = 2.f * q2 + (q2 == 0.f ? RR_MINVAL : 0.f);
  }
}

/* The three candidates of a line-search iteration, one per GROUP OF 8 LANES (the fourth group idles): a group walks all
 * active rows with stride 8 for its own step size, so the reduction is 3 sums over 8 lanes (9 shuffles in 3 dependent rounds)
 * plus one round that hands every lane the three finished points, instead of 9 sums over 32 lanes (45 shuffles in 5 rounds).
 * The line search is the longest serial chain of a solver iteration, and shuffles share the one-per-clock LSU port of the SM
 * with every other warp of the CTA. */
template <int NS>
RR_DEV void ls_eval3(Ctx<NS> &c, int nra, const float (&alpha)[3], float g0, float g1, float g2, LSPoint (&out)[3]) {
  const int grp = c.lane >> 3, sub = c.lane & 7;
  const float al = grp == 0 ? alpha[0] : (grp == 1 ? alpha[1] : alpha[2]);
  float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll 2
  for (int r = sub; r < nra; r += 8) {
    const float ja = c.row_Jaref[r], jv = c.row_jv[r], D = c.row_D[r];
    const float jaD = ja * D, jvD = jv * D;
    if (ja + al * jv < 0.f) { q0 += 0.5f * ja * jaD; q1 += jv * jaD; q2 += 0.5f * jv * jvD; }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    q0 += __shfl_xor_sync(RR_FULL, q0, o);
    q1 += __shfl_xor_sync(RR_FULL, q1, o);
    q2 += __shfl_xor_sync(RR_FULL, q2, o);
  }
  q0 += g0; q1 += g1; q2 += g2;
  const float cost = al * al * q2 + al * q1 + q0, d0 = 2.f * al * q2 + q1, d1 = 2.f * q2 + (q2 == 0.f ? RR_MINVAL : 0.f);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    out[k].alpha = alpha[k];
    out[k].cost = __shfl_sync(RR_FULL, cost, 8 * k);
    out[k].d0 = __shfl_sync(RR_FULL, d0, 8 * k);
    out[k].d1 = __shfl_sync(RR_FULL, d1, 8 * k);
  }
}

/* Given qacc (regs): Jaref = J qacc - aref (rows, smem).  Ma = M qacc comes from the caller: M qacc_warmstart was formed
 * before M was factorised in place, and M qacc_smooth = qfrc_smooth (qacc_smooth solves exactly that system; MJX
 * multiplies it out again, which differs by solve round-off only). */
template <int NS>
RR_DEV void ctx_init(Ctx<NS> &c, const float (&qacc)[NS]) {
  __syncwarp();
  vstore<NS>(c, qacc, c.vbuf);
  __syncwarp();
  mul_j<NS>(c, c.row_Jaref);
  int nra = c.nla + 4 * c.nca;
  for (int r = c.lane; r < nra; r += 32) c.row_Jaref[r] -= c.row_aref[r];
  __syncwarp();
}

/* constraint + Gauss cost at (qacc, Ma, Jaref); with_force also stores efc_force in row_jv */
template <int NS>
RR_DEV float constraint_cost(Ctx<NS> &c, const float (&qacc)[NS], const float (&Ma)[NS], float &gauss, bool with_force) {
  int nra = c.nla + 4 * c.nca;
  float cost = 0.f;
  for (int r = c.lane; r < nra; r += 32) {
    float ja = c.row_Jaref[r], D = c.row_D[r];
    bool act = ja < 0.f;
    if (with_force) c.row_jv[r] = act ? D * -ja : 0.f;
    if (act) cost += D * ja * ja;
  }
  float g = 0.f;
  RR_FOR_S g += (Ma[s] - c.qfrc_smooth[s]) * (qacc[s] - c.qacc_smooth[s]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cost += __shfl_xor_sync(RR_FULL, cost, o);
    g += __shfl_xor_sync(RR_FULL, g, o);
  }
  gauss = 0.5f * g;
  return 0.5f * cost + gauss;
}

/* solver.solve (CG).  The loop is arranged so that each heavy routine (mul_m, mul_j, mul_jt, solve_ld) has one call
 * site: the kernel is instruction-cache bound otherwise (ncu: stall_no_instruction). */
template <int NS>
RR_DEV void solve_constraints(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  const int nra = c.nla + 4 * c.nca;
  const float nvf = (float)(m.nv > 1 ? m.nv : 1);
  const float scale = 1.f / (m.meaninertia * nvf);
  float Ma[NS], grad[NS], Mgrad[NS], search[NS], mv[NS];
  float gauss = 0.f, cost = INFINITY, prev_cost = INFINITY, beta = 0.f, gg = 0.f;
  RR_FOR_S mv[s] = 0.f;
  int niter = 0;
  bool first = true, finished = false;
  if (nra == 0) {
    /* no active row: the cost is the Gauss term alone, minimised (= 0) by qacc_smooth, which therefore wins the warm-start
     * comparison and has zero gradient -- MJX leaves its loop before the first iteration with exactly this state */
    RR_FOR_S { c.qacc[s] = c.qacc_smooth[s]; c.qfrc_constraint[s] = 0.f; }
    cost = 0.f;
    finished = true;
  } else
  /* warm start: keep whichever of qacc_smooth / qacc_warmstart has the lower cost (candidate 2 = back to smooth) */
  {
    float cs = 0.f, cw = 0.f, g;
    for (int cand = 0; cand < 3; cand++) {
      if (cand == 2 && cw < cs) break;
      RR_FOR_S {
        c.qacc[s] = (cand == 1) ? c.warm[s] : c.qacc_smooth[s];
        Ma[s] = (cand == 1) ? c.ma_warm[s] : c.qfrc_smooth[s];
      }
      ctx_init<NS>(c, c.qacc);
      if (cand == 2) break;
      float cst = constraint_cost<NS>(c, c.qacc, Ma, g, false);
      if (cand == 0) cs = cst; else cw = cst;
    }
  }
  prof<NS>(c, RR_PROF_SOLVE_INIT);
  /* fixed trip count (iterations + 1 gradient updates at most) so that the CTA-wide rendezvous below stays matched
   * across warps; a warp whose solve has converged idles through the remaining trips */
#pragma unroll 1
  for (int trip = 0; trip <= m.iterations; trip++) {
    RR_CTA_SYNC_AT(3);
    float prev_grad[NS], prev_Mgrad[NS];
    if (!finished && !first) {
      if (m.iterations != 1) {
        float improvement = (prev_cost - cost) * scale;
        float gradient = RR_SQRT(gg) * scale; /* gg = grad . grad, reduced together with the Polak-Ribiere sums */
        if (niter >= m.iterations || improvement < m.tolerance || gradient < m.tolerance) finished = true;
      } else if (niter >= 1) {
        finished = true;
      }
    }
    if (!finished && !first) {
      /* ---- linesearch ---- */
      /* mv = M search.  search = -Mgrad + beta search_prev with M Mgrad = grad (Mgrad is the LD solve of grad), so
       * mv = -grad + beta mv_prev: the same vector MJX gets from mul_m(search), without the product. */
#if RR_MV_RECURRENCE
      RR_FOR_S mv[s] = -grad[s] + beta * mv[s];
      __syncwarp();
      vstore<NS>(c, search, c.vbuf);
      __syncwarp();
#else
      mul_m<NS>(c, mv, search); /* leaves search staged in vbuf */
#endif
      prof<NS>(c, RR_PROF_LS_PRE);
#pragma unroll 1
      for (int rep = 0; rep <= RR_DUP_MULJ; rep++) mul_j<NS>(c, c.row_jv);
      prof<NS>(c, RR_PROF_LS_MULJ);
      float g0 = gauss, g1 = 0.f, g2 = 0.f, ss = 0.f; /* one reduction for the Gauss quadratic and |search|^2 */
      RR_FOR_S { g1 += search[s] * (Ma[s] - c.qfrc_smooth[s]); g2 += search[s] * mv[s]; ss += search[s] * search[s]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        g1 += __shfl_xor_sync(RR_FULL, g1, o);
        g2 += __shfl_xor_sync(RR_FULL, g2, o);
        ss += __shfl_xor_sync(RR_FULL, ss, o);
      }
      g2 *= 0.5f;
      const float gtol = m.tolerance * m.ls_tolerance * (RR_SQRT(ss) * m.meaninertia * nvf);
      LSPoint p0, lo, hi;
      {
        float a1[1] = {0.f};
        LSPoint r1[1];
        for (int pass = 0; pass < 2; pass++) {
          ls_eval<NS, 1>(c, nra, a1, g0, g1, g2, r1);
          if (pass == 0) { p0 = r1[0]; a1[0] = p0.alpha - p0.d0 * RR_RCP(p0.d1); }
        }
        lo = r1[0];
      }
      if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
      prof<NS>(c, RR_PROF_LS_EVAL2);
      bool swap = true;
      int ls_iter = 0;
      for (;;) {
        bool done = ls_iter >= m.ls_iterations;
        done |= !swap;
        done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
        done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
        if (done) break;
        float a3[3] = {lo.alpha - lo.d0 * RR_RCP(lo.d1), hi.alpha - hi.d0 * RR_RCP(hi.d1), 0.5f * (lo.alpha + hi.alpha)};
        LSPoint r3[3];
#pragma unroll 1
        for (int rep = 0; rep <= RR_DUP_LS; rep++) ls_eval3<NS>(c, nra, a3, g0, g1, g2, r3);
        LSPoint lo_next = r3[0], hi_next = r3[1], mid = r3[2];
        bool swap_lo_next = (lo.d0 > 0.f) || (lo.d0 < lo_next.d0);
        if (swap_lo_next) lo = lo_next;
        bool swap_lo_mid = (mid.d0 < 0.f) && (lo.d0 < mid.d0);
        if (swap_lo_mid) lo = mid;
        bool swap_hi_next = (hi.d0 < 0.f) || (hi.d0 > hi_next.d0);
        if (swap_hi_next) hi = hi_next;
        bool swap_hi_mid = (mid.d0 > 0.f) && (hi.d0 > mid.d0);
        if (swap_hi_mid) hi = mid;
        swap = swap_lo_next | swap_lo_mid | swap_hi_next | swap_hi_mid;
        ls_iter++;
      }
      prof<NS>(c, RR_PROF_LS_LOOP);
      bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
      float alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
      if (improved) {
        RR_FOR_S { c.qacc[s] += search[s] * alpha; Ma[s] += mv[s] * alpha; }
        for (int r = c.lane; r < nra; r += 32) c.row_Jaref[r] += c.row_jv[r] * alpha;
        __syncwarp();
      }
      RR_FOR_S { prev_grad[s] = grad[s]; prev_Mgrad[s] = Mgrad[s]; }
      prof<NS>(c, RR_PROF_SOLVE_LS);
    }
    RR_CTA_SYNC_AT(4); /* the line searches end at different times; realign before the solve */
    if (finished) continue;
    /* ---- _update_constraint + _update_gradient ---- */
    prev_cost = cost;
    prof<NS>(c, RR_PROF_SOLVE_UPD);
    cost = constraint_cost<NS>(c, c.qacc, Ma, gauss, true);
    prof<NS>(c, RR_PROF_COST);
#pragma unroll 1
    for (int rep = 0; rep <= RR_DUP_MULJT; rep++) mul_jt<NS>(c, c.row_jv, c.qfrc_constraint);
    prof<NS>(c, RR_PROF_CRB); /* profiling bucket "crb" = constraint_cost + J' f inside the solver */
    RR_FOR_S { grad[s] = Ma[s] - c.qfrc_smooth[s] - c.qfrc_constraint[s]; }
    if (trip == m.iterations && !first) {
      /* the last gradient update the iteration limit allows: M^-1 grad and the next search direction would never be used
       * (MJX computes and discards them); qfrc_constraint above is what the integrator needs */
      niter++;
      continue;
    }
#pragma unroll 1
    for (int rep = 0; rep <= RR_DUP_SOLVE; rep++) {
      RR_FOR_S Mgrad[s] = grad[s];
      solve_ld<NS>(c, Mgrad, c.LD, c.dinv);
    }
    prof<NS>(c, RR_PROF_VEL); /* profiling bucket "com_vel" = the M^-1 grad solve inside the solver */
    {
      float num = 0.f, den = 0.f;
      gg = 0.f;
      RR_FOR_S { gg += grad[s] * grad[s]; }
      if (!first) RR_FOR_S { num += grad[s] * (Mgrad[s] - prev_Mgrad[s]); den += prev_grad[s] * prev_Mgrad[s]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(RR_FULL, num, o);
        den += __shfl_xor_sync(RR_FULL, den, o);
        gg += __shfl_xor_sync(RR_FULL, gg, o);
      }
      if (first) {
        RR_FOR_S search[s] = -Mgrad[s];
        first = false;
      } else {
        beta = fmaxf(0.f, num * RR_RCP(fmaxf(RR_MINVAL, den)));
        RR_FOR_S search[s] = -Mgrad[s] + beta * search[s];
        niter++;
      }
    }
    prof<NS>(c, RR_PROF_SOLVE_UPD);
  }
  c.niter = niter;
  c.niter_total += niter;
  RR_FOR_S c.warm[s] = c.qacc[s];
  if (RR_WITH_DEBUG && c.dbg) {
    float *dF = c.dbg + dbg_offset(m, RR_DBG_EFC_FORCE), *dS = c.dbg + dbg_offset(m, RR_DBG_SCALARS);
    for (int i = c.lane; i < m.nefc; i += 32) dF[i] = 0.f;
    __syncwarp();
    for (int r = c.lane; r < nra; r += 32) dF[c.row_id[r] & 0xffff] = c.row_jv[r];
    if (c.lane == 0) { dS[0] = (float)niter; dS[1] = (float)c.nla; dS[2] = (float)c.nca; dS[3] = cost; }
    dbg_vec<NS>(c, RR_DBG_QACC, c.qacc);
    dbg_vec<NS>(c, RR_DBG_QFRC_CONSTRAINT, c.qfrc_constraint);
    __syncwarp();
  }
}

/* The observation slices cinert[1:], cvel[1:], qfrc_actuator (Rodent_Env_Brax.py:152-156) and the optional raw views are
 * stored to HBM right after the smooth-force stage of the last forward pass: their shared memory is recycled by the
 * constraint phase.  (brax reports them from the forward pass at the pre-integration qpos, which is this one.) */
template <int NS>
RR_DEV void forward_outputs(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  const RRStepArgs &a = c.a;
  const size_t e = (size_t)c.env;
  const int lane = c.lane, nb1 = m.nbody - 1;
  if (lane < 4 && m.nbody > 1) c.xq1[lane] = c.xquat[4 + lane];
  __syncwarp();
  if (!c.live) return;
  if (a.obs) {
    const int obs_dim = m.nq + m.nv + 16 * nb1 + m.nv + 3;
    float *o = a.obs + e * obs_dim + m.nq + m.nv;
#pragma unroll 1
    for (int i = lane; i < 10 * nb1; i += 32) o[i] = c.cinert[10 + i];
    o += 10 * nb1;
#pragma unroll 1
    for (int i = lane; i < 6 * nb1; i += 32) o[i] = c.cvel[6 + i];
    o += 6 * nb1;
#pragma unroll 1
    for (int i = lane; i < m.nv; i += 32) o[i] = c.qfrc_act[i];
  }
  if (a.xpos) for (int i = lane; i < 3 * m.nbody; i += 32) a.xpos[e * 3 * m.nbody + i] = c.xpos[i];
  if (a.xquat) for (int i = lane; i < 4 * m.nbody; i += 32) a.xquat[e * 4 * m.nbody + i] = c.xquat[i];
  if (a.subtree_com) for (int i = lane; i < 3 * m.nroot; i += 32) a.subtree_com[e * 3 * m.nroot + i] = c.com[i];
  if (a.qfrc_actuator) for (int i = lane; i < m.nv; i += 32) a.qfrc_actuator[e * m.nv + i] = c.qfrc_act[i];
  if (a.cinert) for (int i = lane; i < 10 * m.nbody; i += 32) a.cinert[e * 10 * m.nbody + i] = c.cinert[i];
  if (a.cvel) for (int i = lane; i < 6 * m.nbody; i += 32) a.cvel[e * 6 * m.nbody + i] = c.cvel[i];
}

/* ------------------------------------------------------------------------------------------ forward / euler */
/* One mjx.step: forward (B.1-B.7) then, if `integrate`, the implicit-damping Euler update (B.8).  The two
 * factorisations (M for the solver, M + dt diag(damping) for Euler) and their solves share one call site. */
template <int NS>
RR_DEV void substep(Ctx<NS> &c, bool integrate, float &time, int sub) {
  const RRModelDev &m = c.m;
  const float dt = m.timestep;
  RR_CTA_SYNC_IF(RR_SYNC_LEVEL >= 1 && sub % RR_SYNC_PERIOD == 0);
  __syncwarp();
  prof<NS>(c, RR_PROF_WAIT);
#pragma unroll 1
  for (int rep = 0; rep <= RR_DUP_KIN; rep++) kinematics<NS>(c);
  prof<NS>(c, RR_PROF_FK);
  com_pos<NS>(c);
  prof<NS>(c, RR_PROF_COM);
  dbg_copy<NS>(c, RR_DBG_XPOS, c.xpos, 3 * m.nbody);
  dbg_copy<NS>(c, RR_DBG_XQUAT, c.xquat, 4 * m.nbody);
  dbg_copy<NS>(c, RR_DBG_COM, c.com, 3 * m.nroot);
  dbg_copy<NS>(c, RR_DBG_CINERT, c.cinert, 10 * m.nbody);
  dbg_copy<NS>(c, RR_DBG_CDOF, c.cdof, 6 * m.nv);
#pragma unroll 1
  for (int rep = 0; rep <= RR_DUP_CRB; rep++) crb_and_mass_matrix<NS>(c);
  prof<NS>(c, RR_PROF_QM);
  {
    float qfrc_bias[NS];
#pragma unroll 1
    for (int rep = 0; rep <= RR_DUP_RNE; rep++) com_vel_and_rne<NS>(c, qfrc_bias);
    prof<NS>(c, RR_PROF_RNE);
    dbg_copy<NS>(c, RR_DBG_CVEL, c.cvel, 6 * m.nbody);
    smooth_forces<NS>(c, qfrc_bias);
    prof<NS>(c, RR_PROF_SMOOTH);
  }
  if (c.last_substep) forward_outputs<NS>(c);
  dbg_copy<NS>(c, RR_DBG_M, c.M, m.nM);
  mul_m_tree<NS>(c, c.ma_warm, c.warm); /* the only product with M the solver needs (see ctx_init) */
  prof<NS>(c, RR_PROF_MULM);
  RR_CTA_SYNC_IF(RR_SYNC_LEVEL >= 2 && RR_SYNC_FACTOR);
  __syncwarp();
  prof<NS>(c, RR_PROF_WAIT2);
#if RR_FACTOR_RL
  factor2_rl<NS>(c, dt);
#else
  factor2<NS>(c, dt);
#endif
  prof<NS>(c, RR_PROF_FACTOR2);
  for (int pass = 0; pass < 2; pass++) {
    float x[NS];
    RR_FOR_S x[s] = pass ? c.qfrc_smooth[s] + c.qfrc_constraint[s] : c.qfrc_smooth[s];
    {
      float dv[NS]; /* one call site for both factors (code size) */
      RR_FOR_S dv[s] = pass ? c.dinv2[s] : c.dinv[s];
      solve_ld<NS>(c, x, pass ? c.M : c.LD, dv);
    }
    prof<NS>(c, pass ? RR_PROF_EULER : RR_PROF_FACTOR);
    if (pass == 0) {
      dbg_copy<NS>(c, RR_DBG_LD, c.LD, m.nM);
      RR_FOR_S c.qacc_smooth[s] = x[s];
      dbg_vec<NS>(c, RR_DBG_QACC_SMOOTH, c.qacc_smooth);
      if (m.nefc == 0) {
        RR_FOR_S { c.qacc[s] = c.qacc_smooth[s]; c.qfrc_constraint[s] = 0.f; }
      } else {
        RR_CTA_SYNC_IF(RR_SYNC_LEVEL >= 2 && RR_SYNC_COLLIDE);
        collision<NS>(c);
        prof<NS>(c, RR_PROF_COLLIDE);
        make_constraint<NS>(c);
        prof<NS>(c, RR_PROF_CONSTRAINT);
        solve_constraints<NS>(c);
      }
      if (!integrate) break;
    } else {
      /* _advance: act, qvel, qpos (semi-implicit: positions use the new velocities) */
      for (int u = c.lane; u < m.nu; u += 32) {
        if (RI(act_dyntype, u) != 0) {
          int aa = RI(act_actadr, u);
          c.act[aa] += c.actdot[aa] * dt;
        }
      }
      RR_FOR_S { int i = c.lane + 32 * s; if (i < m.nv) c.qvel[i] += x[s] * dt; }
      __syncwarp();
      for (int j = c.lane; j < m.njnt; j += 32) {
        int qadr = RI(jnt_qposadr, j), da = RI(jnt_dofadr, j);
        if (RI(jnt_type, j) == RR_JNT_FREE) {
#pragma unroll
          for (int k = 0; k < 3; k++) c.qpos[qadr + k] += dt * c.qvel[da + k];
          float v[3] = {c.qvel[da + 3], c.qvel[da + 4], c.qvel[da + 5]}, q[4], qr[4], q2[4];
          float nrm = normalize3(v);
          axis_angle_quat(qr, v, dt * nrm);
#pragma unroll
          for (int k = 0; k < 4; k++) q[k] = c.qpos[qadr + 3 + k];
          quat_mul(q2, q, qr);
          normalize4(q2);
#pragma unroll
          for (int k = 0; k < 4; k++) c.qpos[qadr + 3 + k] = q2[k];
        } else {
          c.qpos[qadr] += dt * c.qvel[da];
        }
      }
      __syncwarp();
      time += dt;
      prof<NS>(c, RR_PROF_EULER);
    }
  }
}

/* Padding pass of a persistent warp (its CTA has fewer environments than warps in this pass): run exactly the rendezvous
 * sequence of substep() / solve_constraints() and nothing else, so the live warps of the CTA get the issue slots. */
template <int NS>
RR_DEV void substep_idle(Ctx<NS> &c, bool integrate, int sub) {
  const RRModelDev &m = c.m;
  RR_CTA_SYNC_IF(RR_SYNC_LEVEL >= 1 && sub % RR_SYNC_PERIOD == 0);
  RR_CTA_SYNC_IF(RR_SYNC_LEVEL >= 2 && RR_SYNC_FACTOR); /* before the factorisations */
  if (m.nefc != 0) {
    RR_CTA_SYNC_IF(RR_SYNC_LEVEL >= 2 && RR_SYNC_COLLIDE); /* before the collision phase */
#pragma unroll 1
    for (int trip = 0; trip <= m.iterations; trip++) { RR_CTA_SYNC_AT(3); RR_CTA_SYNC_AT(4); }
  }
  (void)integrate;
}

/* ------------------------------------------------------------------------------------------ one environment */
template <int NS>
RR_DEV void env_run(const RRModelDev &m, const RRStepArgs &a, int env_in, int slot, float *sm, const int32_t *ti, const float *tf,
                    int lane) {
  const int env = env_in < a.B ? env_in : a.B - 1;
  Ctx<NS> c(m, a, env, slot, sm, ti, tf, lane);
  c.live = env_in < a.B;
  if (!c.live) c.dbg = nullptr;
  if (RR_WITH_DEBUG && a.prof) {
    if (lane < RR_NPROF) c.prof_acc[lane] = 0.f;
    __syncwarp();
    c.tprev = RR_CLOCK();
  }
  const size_t e = (size_t)env;
  /* ---- load state ---- */
#pragma unroll 1
  for (int i = lane; i < m.nq; i += 32) c.qpos[i] = a.in_qpos[e * m.nq + i];
  for (int i = lane; i < m.nv; i += 32) c.qvel[i] = a.in_qvel[e * m.nv + i];
  for (int i = lane; i < m.na; i += 32) { c.act[i] = a.in_act[e * m.na + i]; c.actdot[i] = 0.f; }
#pragma unroll 1
  for (int i = lane; i < m.nu; i += 32) c.ctrl[i] = a.action ? a.action[e * m.nu + i] : 0.f;
  RR_FOR_S { int i = lane + 32 * s; c.warm[s] = i < m.nv ? a.in_warm[e * m.nv + i] : 0.f; }
  float time = a.in_time ? a.in_time[e] : 0.f;
  /* Brax AutoResetWrapper: steps are zeroed where the previous step ended an episode */
  float steps = 0.f;
  if (a.wrap && a.mode == RR_MODE_STEP) steps = (a.in_done[e] != 0.f) ? 0.f : a.in_steps[e];
  __syncwarp();
  prof<NS>(c, RR_PROF_LOAD);
  /* ---- physics ---- */
  {
    const int nrun = a.mode == RR_MODE_INIT ? 1 : a.nsub;
    if (!c.live) {
      for (int sub = 0; sub < nrun; sub++) substep_idle<NS>(c, a.mode != RR_MODE_INIT, sub);
      return;
    }
    for (int sub = 0; sub < nrun; sub++) {
      c.last_substep = sub == nrun - 1;
      substep<NS>(c, a.mode != RR_MODE_INIT, time, sub);
    }
  }
  __syncwarp();
  if (!c.live) return;
  /* ---- run-task epilogue (Rodent_Env_Brax.py:98-162) ---- */
  const RRTask &t = a.task;
  int cf = a.in_cur_frame ? a.in_cur_frame[e] : 0;
  float reward = 0.f, done = 0.f, pos_reward = 0.f, quadctrl = 0.f, alive = 0.f;
  int cf_new = cf;
  if (a.mode == RR_MODE_STEP) {
    cf_new = cf + 1;
    int ti = cf < 0 ? 0 : (cf >= t.track_len ? t.track_len - 1 : cf);
    float dx = c.qpos[0] - t.track_pos[3 * ti], dy = c.qpos[1] - t.track_pos[3 * ti + 1], dz = c.qpos[2] - t.track_pos[3 * ti + 2];
    pos_reward = expf(-100.f * sqrtf(dx * dx + dy * dy + dz * dz));
    float z = c.qpos[2];
    float healthy = z < t.healthy_z_lo ? 0.f : 1.f;
    healthy = z > t.healthy_z_hi ? 0.f : healthy;
    alive = t.terminate_when_unhealthy ? t.healthy_reward : t.healthy_reward * healthy;
    float sq = 0.f;
    for (int i = lane; i < m.nu; i += 32) sq += c.ctrl[i] * c.ctrl[i];
    float ctrl_cost = t.ctrl_cost_weight * warp_sum(sq);
    quadctrl = -ctrl_cost;
    reward = pos_reward + alive - ctrl_cost;
    done = t.terminate_when_unhealthy ? 1.f - healthy : 0.f;
  }
  /* EpisodeWrapper */
  float trunc = 0.f;
  if (a.wrap && a.mode == RR_MODE_STEP) {
    steps += 1.f;
    if (steps >= (float)a.episode_length) { trunc = 1.f - done; done = 1.f; }
  }
  const bool restore = a.wrap && a.mode == RR_MODE_STEP && done != 0.f; /* AutoResetWrapper */
  /* ---- state write-back ---- */
  if (restore) {
#pragma unroll 1
    for (int i = lane; i < m.nq; i += 32) a.qpos[e * m.nq + i] = a.first_qpos[e * m.nq + i];
    for (int i = lane; i < m.nv; i += 32) a.qvel[e * m.nv + i] = a.first_qvel[e * m.nv + i];
#pragma unroll 1
    for (int i = lane; i < m.na; i += 32) a.act[e * m.na + i] = a.first_act[e * m.na + i];
    for (int i = lane; i < m.nv; i += 32) a.warm[e * m.nv + i] = a.first_warm[e * m.nv + i];
    if (a.time && lane == 0) a.time[e] = a.first_time ? a.first_time[e] : 0.f;
  } else {
#pragma unroll 1
    for (int i = lane; i < m.nq; i += 32) a.qpos[e * m.nq + i] = c.qpos[i];
    for (int i = lane; i < m.nv; i += 32) a.qvel[e * m.nv + i] = c.qvel[i];
#pragma unroll 1
    for (int i = lane; i < m.na; i += 32) a.act[e * m.na + i] = c.act[i];
    RR_FOR_S { int i = lane + 32 * s; if (i < m.nv) a.warm[e * m.nv + i] = c.warm[s]; }
    if (a.time && lane == 0) a.time[e] = time;
  }
  /* ---- observation (Rodent_Env_Brax.py:138-162) ---- */
  if (a.obs) {
    const int nb1 = m.nbody - 1;
    const int obs_dim = m.nq + m.nv + 16 * nb1 + m.nv + 3;
    float *o = a.obs + e * obs_dim;
    if (restore) {
      const float *fo = a.first_obs + e * obs_dim;
#pragma unroll 1
      for (int i = lane; i < obs_dim; i += 32) o[i] = fo[i];
    } else {
#pragma unroll 1
      for (int i = lane; i < m.nq; i += 32) o[i] = c.qpos[i];
      o += m.nq;
#pragma unroll 1
      for (int i = lane; i < m.nv; i += 32) o[i] = c.qvel[i];
      o += m.nv + 16 * nb1 + m.nv; /* cinert / cvel / qfrc_actuator were stored by forward_outputs */
      if (lane < 3) {
        int ti = cf_new + 1;
        ti = ti < 0 ? 0 : (ti >= t.track_len ? t.track_len - 1 : ti);
        float v[3] = {t.track_pos[3 * ti] - c.qpos[0], t.track_pos[3 * ti + 1] - c.qpos[1], t.track_pos[3 * ti + 2] - c.qpos[2]};
        float R[9], xq[4] = {c.xq1[0], c.xq1[1], c.xq1[2], c.xq1[3]};
        quat_to_mat(R, xq);
        o[lane] = R[3 * lane] * v[0] + R[3 * lane + 1] * v[1] + R[3 * lane + 2] * v[2]; /* xmat[1] @ v (not transposed) */
      }
    }
  }
  if (lane == 0) {
    if (a.cur_frame) a.cur_frame[e] = cf_new;
    if (a.reward) a.reward[e] = reward;
    if (a.done) a.done[e] = done;
    if (a.metrics) { a.metrics[3 * e] = pos_reward; a.metrics[3 * e + 1] = quadctrl; a.metrics[3 * e + 2] = alive; }
    if (a.wrap) { a.steps[e] = a.mode == RR_MODE_INIT ? 0.f : steps; a.truncation[e] = trunc; }
    if (a.niter) a.niter[e] = c.niter;
    if (a.work) a.work[e] = (float)c.niter_total;
  }
  /* optional raw outputs that survive the solver phase */
  if (a.contact_dist && m.nefc) for (int i = lane; i < m.ncon; i += 32) a.contact_dist[e * m.ncon + i] = c.con_dist[i];
  if (a.contact_pos && m.nefc)
    for (int i = lane; i < 3 * m.ncon; i += 32) {
      const int cc = i / 3, k = i - 3 * cc;
      a.contact_pos[e * 3 * m.ncon + i] = c.cab[12 * cc + 9 + k] + c.com[3 * RI(body_rootslot, RI(pair_body, RI(con_pair, cc))) + k];
    }
  if (a.contact_frame && m.nefc)
    for (int i = lane; i < 9 * m.ncon; i += 32) a.contact_frame[e * 9 * m.ncon + i] = c.cab[12 * (i / 9) + i % 9];
  if (a.qacc) RR_FOR_S { int i = lane + 32 * s; if (i < m.nv) a.qacc[e * m.nv + i] = c.qacc[s]; }
  prof<NS>(c, RR_PROF_EPILOGUE);
  prof_flush<NS>(c);
}

}  // namespace RR_NS

