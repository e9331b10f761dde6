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
#ifndef RR_KERNELS_INL_
#define RR_KERNELS_INL_

#include "../../include/rr_model_fields.h"
#include "rr_device.h"

#ifndef RR_DEV
#define RR_DEV __device__ __forceinline__
#define RR_HOSTDEV __host__ __device__ inline
#define RR_DEV_MEMBER __device__ __forceinline__
#define RR_LDG(p) __ldg(p)
#define RR_CLOCK() clock64()
#endif

#define RR_FULL 0xffffffffu
#define RR_MINVAL 1e-15f
#define RR_MINIMP 0.0001f
#define RR_MAXIMP 0.9999f
#define RR_SIGN_BIT 0x40000000

namespace rr {

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
  sincosf(angle * 0.5f, &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
RR_DEV float normalize3(float *v) {
  float n = sqrtf(dot3(v, v));
  float d = n + 1e-6f * (n == 0.f ? 1.f : 0.f);
  v[0] /= d; v[1] /= d; v[2] /= d;
  return n;
}
RR_DEV void normalize4(float *v) {
  float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
  float d = n + 1e-6f * (n == 0.f ? 1.f : 0.f);
  v[0] /= d; v[1] /= d; v[2] /= d; v[3] /= d;
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
RR_DEV void kbi(float timestep, const float *solref, const float *solimp, float pos, float &k, float &b, float &imp) {
  float timeconst = fmaxf(solref[0], 2.f * timestep), dampratio = solref[1];
  float dmin = clampf(solimp[0], RR_MINIMP, RR_MAXIMP), dmax = clampf(solimp[1], RR_MINIMP, RR_MAXIMP);
  float width = fmaxf(solimp[2], RR_MINVAL), mid = clampf(solimp[3], RR_MINIMP, RR_MAXIMP), power = fmaxf(solimp[4], 1.f);
  k = 1.f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = 2.f / (dmax * timeconst);
  if (solref[0] <= 0.f) k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0.f) b = -solref[1] / dmax;
  float x = fabsf(pos) / width;
  float ia = (1.f / powf(mid, power - 1.f)) * powf(x, power);
  float ib = 1.f - (1.f / powf(1.f - mid, power - 1.f)) * powf(1.f - x, power);
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
  float *qpos, *qvel, *act, *ctrl, *actdot, *xpos, *xquat, *com, *cinert, *cdof, *cvel, *M, *LD, *Dinv, *vbuf, *qfrc_act;
  float *crb, *fcrb, *cacc, *cfrc, *con_dist, *con_pos, *con_frame, *con_J, *row_D, *row_aref, *row_Jaref, *row_jv;
  int *row_id, *cact;
  /* per-lane dof metadata: dof i = lane + 32 s */
  int radr[NS], dep[NS], nd[NS];
  float dinv[NS];
  /* register-resident nv-vectors */
  float qfrc_smooth[NS], qacc_smooth[NS], warm[NS], qacc[NS], qfrc_constraint[NS];
  int nla, nca; /* active limit rows, active contacts; rows = nla + 4 nca */
  int niter;
  float *dbg;
  long long tprev;

  RR_DEV_MEMBER Ctx(const RRModelDev &m_, const RRStepArgs &a_, int env_, float *sm, int lane_) : m(m_), a(a_), env(env_), lane(lane_) {
    const RRSmem &s = m.sm;
    qpos = sm + s.qpos; qvel = sm + s.qvel; act = sm + s.act; ctrl = sm + s.ctrl; actdot = sm + s.actdot;
    xpos = sm + s.xpos; xquat = sm + s.xquat; com = sm + s.com; cinert = sm + s.cinert; cdof = sm + s.cdof;
    cvel = sm + s.cvel; M = sm + s.M; LD = sm + s.LD; Dinv = sm + s.Dinv; vbuf = sm + s.vbuf; qfrc_act = sm + s.qfrc_act;
    crb = sm + s.crb; fcrb = sm + s.fcrb; cacc = sm + s.cacc; cfrc = sm + s.cfrc;
    con_dist = sm + s.con_dist; con_pos = sm + s.con_pos; con_frame = sm + s.con_frame; con_J = sm + s.con_J;
    row_D = sm + s.row_D; row_aref = sm + s.row_aref; row_Jaref = sm + s.row_Jaref; row_jv = sm + s.row_jv;
    row_id = (int *)(sm + s.row_id); cact = (int *)(sm + s.cact);
#pragma unroll
    for (int s_ = 0; s_ < NS; s_++) {
      int i = lane + 32 * s_;
      bool v = i < m.nv;
      radr[s_] = v ? RR_LDG(&m.dof_rowadr[i]) : 0;
      dep[s_] = v ? RR_LDG(&m.dof_depth[i]) : 0;
      nd[s_] = v ? RR_LDG(&m.dof_ndesc[i]) : 0;
      dinv[s_] = 0.f;
    }
    nla = nca = 0;
    niter = 0;
    dbg = a.dbg.buf ? a.dbg.buf + (size_t)env * a.dbg.stride : nullptr;
    tprev = 0;
  }
};

#define RR_FOR_S _Pragma("unroll") for (int s = 0; s < NS; s++)

template <int NS>
RR_DEV void prof(Ctx<NS> &c, int id) {
  if (c.a.prof) {
    long long t = RR_CLOCK();
    if (c.lane == 0) c.a.prof[(size_t)c.env * RR_NPROF + id] += t - c.tprev;
    c.tprev = t;
  }
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
  if (!c.dbg) return;
  float *dst = c.dbg + dbg_offset(c.m, field);
  for (int i = c.lane; i < n; i += 32) dst[i] = src[i];
}
template <int NS>
RR_DEV void dbg_vec(Ctx<NS> &c, int field, const float (&x)[NS]) {
  if (!c.dbg) return;
  float *dst = c.dbg + dbg_offset(c.m, field);
  RR_FOR_S { int i = c.lane + 32 * s; if (i < c.m.nv) dst[i] = x[s]; }
}

/* ------------------------------------------------------------------------------------------ kinematics (B.1) */
template <int NS>
RR_DEV void kinematics(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  if (c.lane == 0) {
    c.xpos[0] = c.xpos[1] = c.xpos[2] = 0.f;
    c.xquat[0] = 1.f; c.xquat[1] = c.xquat[2] = c.xquat[3] = 0.f;
    c.cinert[6] = c.cinert[7] = c.cinert[8] = 0.f; /* xipos of world (temp slot) */
  }
  __syncwarp();
  for (int lev = 1; lev < m.nlevel; lev++) {
    int beg = RR_LDG(&m.level_adr[lev]), end = RR_LDG(&m.level_adr[lev + 1]);
    for (int idx = beg + c.lane; idx < end; idx += 32) {
      int b = RR_LDG(&m.level_body[idx]);
      int p = RR_LDG(&m.body_parentid[b]);
      float ppos[3], pquat[4], bp[3], bq[4], pos[3], quat[4], r[3];
#pragma unroll
      for (int k = 0; k < 3; k++) { ppos[k] = c.xpos[3 * p + k]; bp[k] = RR_LDG(&m.body_pos[3 * b + k]); }
#pragma unroll
      for (int k = 0; k < 4; k++) { pquat[k] = c.xquat[4 * p + k]; bq[k] = RR_LDG(&m.body_quat[4 * b + k]); }
      rotq(r, bp, pquat);
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] = ppos[k] + r[k];
      quat_mul(quat, pquat, bq);
      int jadr = RR_LDG(&m.body_jntadr[b]), jnum = RR_LDG(&m.body_jntnum[b]);
      for (int j = jadr; j < jadr + jnum; j++) {
        int qa = RR_LDG(&m.jnt_qposadr[j]), da = RR_LDG(&m.jnt_dofadr[j]);
        if (RR_LDG(&m.jnt_type[j]) == RR_JNT_FREE) {
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
          for (int k = 0; k < 3; k++) { jp[k] = RR_LDG(&m.jnt_pos[3 * j + k]); ja[k] = RR_LDG(&m.jnt_axis[3 * j + k]); }
          rotq(r, jp, quat);
#pragma unroll
          for (int k = 0; k < 3; k++) anchor[k] = r[k] + pos[k];
          rotq(axis, ja, quat);
          axis_angle_quat(qloc, ja, c.qpos[qa] - RR_LDG(&m.qpos0[qa]));
          quat_mul(q2, quat, qloc);
#pragma unroll
          for (int k = 0; k < 4; k++) quat[k] = q2[k];
          rotq(r, jp, quat);
#pragma unroll
          for (int k = 0; k < 3; k++) {
            pos[k] = anchor[k] - r[k];
            c.cdof[6 * da + k] = axis[k];       /* temp: xaxis */
            c.cdof[6 * da + 3 + k] = anchor[k]; /* temp: xanchor */
          }
        }
      }
      float ip[3];
#pragma unroll
      for (int k = 0; k < 3; k++) { c.xpos[3 * b + k] = pos[k]; ip[k] = RR_LDG(&m.body_ipos[3 * b + k]); }
#pragma unroll
      for (int k = 0; k < 4; k++) c.xquat[4 * b + k] = quat[k];
      rotq(r, ip, quat);
#pragma unroll
      for (int k = 0; k < 3; k++) c.cinert[10 * b + 6 + k] = pos[k] + r[k]; /* temp: xipos */
    }
    __syncwarp();
  }
}

/* ------------------------------------------------------------------------------------------ com_pos (B.2) */
template <int NS>
RR_DEV void com_pos(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  for (int r = 0; r < m.nroot; r++) {
    float sx = 0.f, sy = 0.f, sz = 0.f, sm = 0.f;
    for (int b = 1 + c.lane; b < m.nbody; b += 32) {
      if (RR_LDG(&m.body_rootslot[b]) == r) {
        float mass = RR_LDG(&m.body_mass[b]);
        sx += c.cinert[10 * b + 6] * mass; sy += c.cinert[10 * b + 7] * mass; sz += c.cinert[10 * b + 8] * mass;
        sm += mass;
      }
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sm = warp_sum(sm);
    if (c.lane == 0) {
      bool tiny = sm < RR_MINVAL;
      c.com[3 * r + 0] = tiny ? 0.f : sx / sm;
      c.com[3 * r + 1] = tiny ? 0.f : sy / sm;
      c.com[3 * r + 2] = tiny ? 0.f : sz / sm;
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
    for (int k = 0; k < 4; k++) { xq[k] = c.xquat[4 * b + k]; iq[k] = RR_LDG(&m.body_iquat[4 * b + k]); }
    quat_mul(q, xq, iq);
    quat_to_mat(R, q);
    int rs = RR_LDG(&m.body_rootslot[b]);
    float mb = RR_LDG(&m.body_mass[b]);
#pragma unroll
    for (int k = 0; k < 3; k++) { I[k] = RR_LDG(&m.body_inertia[3 * b + k]); off[k] = ci[6 + k] - c.com[3 * rs + k]; }
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
    int b = RR_LDG(&m.jnt_bodyid[j]), da = RR_LDG(&m.jnt_dofadr[j]);
    int rs = RR_LDG(&m.body_rootslot[b]);
    float *cd = c.cdof + 6 * da;
    if (RR_LDG(&m.jnt_type[j]) == RR_JNT_FREE) {
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
  for (int b = m.nbody - 1; b > 0; b--) {
    int p = RR_LDG(&m.body_parentid[b]);
    if (c.lane < 10 && p > 0) c.crb[10 * p + c.lane] += c.crb[10 * b + c.lane];
    __syncwarp();
  }
  for (int i = c.lane; i < m.nv; i += 32) {
    float f[6], cr[10], cd[6];
    int b = RR_LDG(&m.dof_bodyid[i]);
#pragma unroll
    for (int k = 0; k < 10; k++) cr[k] = c.crb[10 * b + k];
#pragma unroll
    for (int k = 0; k < 6; k++) cd[k] = c.cdof[6 * i + k];
    inert_mul(f, cr, cd);
#pragma unroll
    for (int k = 0; k < 6; k++) c.fcrb[6 * i + k] = f[k];
  }
  __syncwarp();
  for (int e = c.lane; e < m.nM; e += 32) {
    int i = RR_LDG(&m.M_rowid[e]), j = RR_LDG(&m.M_colind[e]);
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) v += c.cdof[6 * j + k] * c.fcrb[6 * i + k];
    if (i == j) v += RR_LDG(&m.dof_armature[i]);
    c.M[e] = v;
  }
  __syncwarp();
}

/* Tree-sparse LDL' of (M + diag_add) into LD / Dinv (MuJoCo mj_factorM order: leaves to root).  Equivalent to the
 * dense Cholesky MJX runs (jax.scipy cho_factor) up to rounding. */
template <int NS>
RR_DEV void factor(Ctx<NS> &c, float diag_scale) {
  const RRModelDev &m = c.m;
  for (int e = c.lane; e < m.nM; e += 32) {
    float v = c.M[e];
    int i = RR_LDG(&m.M_rowid[e]);
    if (diag_scale != 0.f && i == RR_LDG(&m.M_colind[e])) v += diag_scale * RR_LDG(&m.dof_damping[i]);
    c.LD[e] = v;
  }
  __syncwarp();
  for (int k = m.nv - 1; k > 0; k--) {
    int mk = RR_LDG(&m.dof_depth[k]);
    if (mk == 0) continue;
    int adr = RR_LDG(&m.dof_rowadr[k]);
    float dk = c.LD[adr + mk];
    for (int t = c.lane; t < mk; t += 32) {
      int i = RR_LDG(&m.M_colind[adr + t]);
      int ra = RR_LDG(&m.dof_rowadr[i]);
      float tmp = c.LD[adr + t] / dk;
      for (int s = 0; s <= t; s++) c.LD[ra + s] -= c.LD[adr + s] * tmp;
    }
    __syncwarp();
    for (int t = c.lane; t < mk; t += 32) c.LD[adr + t] = c.LD[adr + t] / dk;
    __syncwarp();
  }
  for (int i = c.lane; i < m.nv; i += 32) c.Dinv[i] = 1.f / c.LD[RR_LDG(&m.dof_rowadr[i]) + RR_LDG(&m.dof_depth[i])];
  __syncwarp();
  RR_FOR_S { int i = c.lane + 32 * s; c.dinv[s] = i < m.nv ? c.Dinv[i] : 0.f; }
}

/* x <- (L D L')^-1 x with x distributed over lanes (dof i = lane + 32 s); pure register / shuffle solve. */
template <int NS>
RR_DEV void solve_ld(Ctx<NS> &c, float (&x)[NS]) {
  const RRModelDev &m = c.m;
  for (int i = m.nv - 1; i > 0; i--) {
    float xi = __shfl_sync(RR_FULL, vselect<NS>(x, i >> 5), i & 31);
    int adr = RR_LDG(&m.dof_rowadr[i]);
    RR_FOR_S {
      int j = c.lane + 32 * s;
      if (j < i && i <= j + c.nd[s]) x[s] -= c.LD[adr + c.dep[s]] * xi;
    }
  }
  RR_FOR_S x[s] *= c.dinv[s];
  for (int j = 0; j < m.nv - 1; j++) {
    int ndj = RR_LDG(&m.dof_ndesc[j]);
    if (ndj == 0) continue;
    float xj = __shfl_sync(RR_FULL, vselect<NS>(x, j >> 5), j & 31);
    int depj = RR_LDG(&m.dof_depth[j]);
    RR_FOR_S {
      int i = c.lane + 32 * s;
      if (i > j && i <= j + ndj) x[s] -= c.LD[c.radr[s] + depj] * xj;
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
      for (int t = 0; t <= c.dep[s]; t++) acc += c.M[adr + t] * c.vbuf[RR_LDG(&m.M_colind[adr + t])];
      for (int k = i + 1; k <= i + c.nd[s]; k++) acc += c.M[RR_LDG(&m.dof_rowadr[k]) + c.dep[s]] * c.vbuf[k];
    }
    y[s] = acc;
  }
  __syncwarp();
}

/* ------------------------------------------------------------------------------------------ velocity + rne (B.6) */
template <int NS>
RR_DEV void com_vel_and_rne(Ctx<NS> &c, float (&qfrc_bias)[NS]) {
  const RRModelDev &m = c.m;
  if (c.lane < 6) {
    c.cvel[c.lane] = 0.f;
    c.cacc[c.lane] = c.lane < 3 ? 0.f : -m.gravity[c.lane - 3];
  }
  __syncwarp();
  for (int lev = 1; lev < m.nlevel; lev++) {
    int beg = RR_LDG(&m.level_adr[lev]), end = RR_LDG(&m.level_adr[lev + 1]);
    for (int idx = beg + c.lane; idx < end; idx += 32) {
      int b = RR_LDG(&m.level_body[idx]);
      int p = RR_LDG(&m.body_parentid[b]);
      float cv[6], ca[6];
#pragma unroll
      for (int k = 0; k < 6; k++) { cv[k] = c.cvel[6 * p + k]; ca[k] = c.cacc[6 * p + k]; }
      int jadr = RR_LDG(&m.body_jntadr[b]), jnum = RR_LDG(&m.body_jntnum[b]);
      for (int j = jadr; j < jadr + jnum; j++) {
        int d0 = RR_LDG(&m.jnt_dofadr[j]);
        if (RR_LDG(&m.jnt_type[j]) == RR_JNT_FREE) {
#pragma unroll
          for (int d = 0; d < 3; d++) {
            float qv = c.qvel[d0 + d];
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
            float qv = c.qvel[d0 + 3 + d];
#pragma unroll
            for (int k = 0; k < 6; k++) { cv[k] += c.cdof[6 * (d0 + 3 + d) + k] * qv; ca[k] += cdd[d][k] * qv; }
          }
        } else {
          float cd[6], cdd[6], qv = c.qvel[d0];
#pragma unroll
          for (int k = 0; k < 6; k++) cd[k] = c.cdof[6 * d0 + k];
          motion_cross(cdd, cv, cd);
#pragma unroll
          for (int k = 0; k < 6; k++) { cv[k] += cd[k] * qv; ca[k] += cdd[k] * qv; }
        }
      }
#pragma unroll
      for (int k = 0; k < 6; k++) { c.cvel[6 * b + k] = cv[k]; c.cacc[6 * b + k] = ca[k]; }
    }
    __syncwarp();
  }
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
    for (int k = 0; k < 6; k++) c.cfrc[6 * b + k] = f1[k] + f3[k];
  }
  __syncwarp();
  for (int b = m.nbody - 1; b > 0; b--) {
    int p = RR_LDG(&m.body_parentid[b]);
    if (c.lane < 6 && p > 0) c.cfrc[6 * p + c.lane] += c.cfrc[6 * b + c.lane];
    __syncwarp();
  }
  RR_FOR_S {
    int i = c.lane + 32 * s;
    float v = 0.f;
    if (i < m.nv) {
      int b = RR_LDG(&m.dof_bodyid[i]);
#pragma unroll
      for (int k = 0; k < 6; k++) v += c.cdof[6 * i + k] * c.cfrc[6 * b + k];
    }
    qfrc_bias[s] = v;
  }
}

/* passive + actuation -> qfrc_smooth, qacc_smooth */
template <int NS>
RR_DEV void smooth_forces(Ctx<NS> &c, const float (&qfrc_bias)[NS]) {
  const RRModelDev &m = c.m;
  __syncwarp();
  /* passive springs into vbuf (per joint), dampers added per dof below */
  for (int i = c.lane; i < m.nv; i += 32) { c.vbuf[i] = 0.f; c.qfrc_act[i] = 0.f; }
  __syncwarp();
  for (int j = c.lane; j < m.njnt; j += 32) {
    int qa = RR_LDG(&m.jnt_qposadr[j]), da = RR_LDG(&m.jnt_dofadr[j]);
    float k = RR_LDG(&m.jnt_stiffness[j]);
    if (RR_LDG(&m.jnt_type[j]) == RR_JNT_FREE) {
      /* free-joint spring: translational part only matters when stiffness != 0 (never for <freejoint>) */
#pragma unroll
      for (int d = 0; d < 3; d++) c.vbuf[da + d] = -k * (c.qpos[qa + d] - RR_LDG(&m.qpos_spring[qa + d]));
    } else {
      c.vbuf[da] = -k * (c.qpos[qa] - RR_LDG(&m.qpos_spring[qa]));
    }
  }
  /* actuation (fwd_actuation): filter activation, affine gain / bias, joint transmission */
  for (int u = c.lane; u < m.nu; u += 32) {
    float ctrl = c.ctrl[u];
    if (RR_LDG(&m.act_ctrllimited[u])) ctrl = clampf(ctrl, RR_LDG(&m.act_ctrlrange[2 * u]), RR_LDG(&m.act_ctrlrange[2 * u + 1]));
    int da = RR_LDG(&m.act_dofadr[u]), qa = RR_LDG(&m.act_qposadr[u]);
    float gear = RR_LDG(&m.act_gear[u]);
    float len = gear * c.qpos[qa], vel = gear * c.qvel[da];
    float ctrl_act = ctrl;
    if (RR_LDG(&m.act_dyntype[u]) == 2) {
      int aa = RR_LDG(&m.act_actadr[u]);
      float tau = fmaxf(RR_LDG(&m.act_dynprm[u]), RR_MINVAL);
      c.actdot[aa] = (ctrl - c.act[aa]) / tau;
      ctrl_act = c.act[aa];
    }
    float gain = RR_LDG(&m.act_gainprm[3 * u]);
    if (RR_LDG(&m.act_gaintype[u]) == 1) gain += RR_LDG(&m.act_gainprm[3 * u + 1]) * len + RR_LDG(&m.act_gainprm[3 * u + 2]) * vel;
    float bias = 0.f;
    if (RR_LDG(&m.act_biastype[u]) == 1)
      bias = RR_LDG(&m.act_biasprm[3 * u]) + RR_LDG(&m.act_biasprm[3 * u + 1]) * len + RR_LDG(&m.act_biasprm[3 * u + 2]) * vel;
    float force = gain * ctrl_act + bias;
    if (RR_LDG(&m.act_forcelimited[u])) force = clampf(force, RR_LDG(&m.act_forcerange[2 * u]), RR_LDG(&m.act_forcerange[2 * u + 1]));
    c.qfrc_act[da] = gear * force;
  }
  __syncwarp();
  float passive[NS];
  RR_FOR_S {
    int i = c.lane + 32 * s;
    float pv = 0.f, av = 0.f;
    if (i < m.nv) { pv = c.vbuf[i] - RR_LDG(&m.dof_damping[i]) * c.qvel[i]; av = c.qfrc_act[i]; }
    passive[s] = pv;
    c.qfrc_smooth[s] = pv - qfrc_bias[s] + av;
    c.qacc_smooth[s] = c.qfrc_smooth[s];
  }
  dbg_vec<NS>(c, RR_DBG_QFRC_BIAS, qfrc_bias);
  dbg_vec<NS>(c, RR_DBG_QFRC_PASSIVE, passive);
  dbg_copy<NS>(c, RR_DBG_QFRC_ACTUATOR, c.qfrc_act, m.nv);
  solve_ld<NS>(c, c.qacc_smooth);
  dbg_vec<NS>(c, RR_DBG_QFRC_SMOOTH, c.qfrc_smooth);
  dbg_vec<NS>(c, RR_DBG_QACC_SMOOTH, c.qacc_smooth);
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
    int b = RR_LDG(&m.pair_body[p]), ca = RR_LDG(&m.pair_conadr[p]), fn = RR_LDG(&m.pair_fn[p]);
    float n[3], pp[3], gl[3], gq[4], xq[4], gp[3], r[3], size[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      n[k] = RR_LDG(&m.pair_plane_n[3 * p + k]); pp[k] = RR_LDG(&m.pair_plane_p[3 * p + k]);
      gl[k] = RR_LDG(&m.pair_gpos[3 * p + k]); size[k] = RR_LDG(&m.pair_size[3 * p + k]);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) { xq[k] = c.xquat[4 * b + k]; gq[k] = RR_LDG(&m.pair_gquat[4 * p + k]); }
    rotq(r, gl, xq);
#pragma unroll
    for (int k = 0; k < 3; k++) gp[k] = c.xpos[3 * b + k] + r[k];
    if (fn == RR_PAIR_PLANE_SPHERE) {
      float d[3] = {gp[0] - pp[0], gp[1] - pp[1], gp[2] - pp[2]};
      float dist = dot3(d, n) - size[0];
      c.con_dist[ca] = dist;
#pragma unroll
      for (int k = 0; k < 3; k++) c.con_pos[3 * ca + k] = gp[k] - n[k] * (size[0] + 0.5f * dist);
      make_frame(c.con_frame + 9 * ca, n);
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
          for (int k = 0; k < 3; k++) c.con_pos[3 * (ca + e) + k] = cp[k] - n[k] * (size[0] + 0.5f * dist);
#pragma unroll
          for (int k = 0; k < 9; k++) c.con_frame[9 * (ca + e) + k] = frame[k];
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
        for (int k = 0; k < 3; k++) c.con_pos[3 * ca + k] = wp[k] - n[k] * dist * 0.5f;
        make_frame(c.con_frame + 9 * ca, n);
      }
    }
  }
  __syncwarp();
  dbg_copy<NS>(c, RR_DBG_CON_DIST, c.con_dist, m.ncon);
  dbg_copy<NS>(c, RR_DBG_CON_POS, c.con_pos, 3 * m.ncon);
  dbg_copy<NS>(c, RR_DBG_CON_FRAME, c.con_frame, 9 * m.ncon);
}

/* ------------------------------------------------------------------------------------------ constraint rows (B.5) */
/* rows: jv[r] = J_r . v for the compact active rows; v staged in vbuf by the caller (already synced). */
template <int NS>
RR_DEV void mul_j(Ctx<NS> &c, float *out) {
  const RRModelDev &m = c.m;
  for (int r = c.lane; r < c.nla; r += 32) {
    int id = c.row_id[r];
    float sg = (id & RR_SIGN_BIT) ? -1.f : 1.f;
    out[r] = sg * c.vbuf[RR_LDG(&m.limit_dofadr[id & 0xffff])];
  }
  /* contacts: 3 frame-row dot products per contact into con_dist[3k..] scratch (dist no longer needed) */
  float *a3 = c.con_pos; /* 3 * ncon scratch: con_pos is dead once the Jacobians exist */
  for (int it = c.lane; it < 3 * c.nca; it += 32) {
    int k = it / 3, r3 = it - 3 * k;
    int cc = c.cact[k];
    int p = RR_LDG(&m.con_pair[cc]);
    int ld = RR_LDG(&m.pair_lastdof[p]);
    int len = RR_LDG(&m.dof_depth[ld]) + 1, adr = RR_LDG(&m.dof_rowadr[ld]);
    const float *J = c.con_J + RR_LDG(&m.con_Jadr[cc]) + r3 * len;
    float acc = 0.f;
    for (int t = 0; t < len; t++) acc += J[t] * c.vbuf[RR_LDG(&m.M_colind[adr + t])];
    a3[it] = acc;
  }
  __syncwarp();
  for (int r = c.lane; r < 4 * c.nca; r += 32) {
    int k = r >> 2, q = r & 3;
    float mu = RR_LDG(&m.pair_mu[RR_LDG(&m.con_pair[c.cact[k]])]);
    float f = (q & 1) ? -mu : mu;
    out[c.nla + r] = a3[3 * k] + a3[3 * k + 1 + (q >> 1)] * f;
  }
  __syncwarp();
}

/* qfc = J' f for the compact active rows; f in `frc` (smem rows) */
template <int NS>
RR_DEV void mul_jt(Ctx<NS> &c, const float *frc, float (&qfc)[NS]) {
  const RRModelDev &m = c.m;
  float *g3 = c.con_pos;
  __syncwarp();
  for (int i = c.lane; i < m.nv; i += 32) c.vbuf[i] = 0.f;
  __syncwarp();
  for (int r = c.lane; r < c.nla; r += 32) {
    int id = c.row_id[r];
    float sg = (id & RR_SIGN_BIT) ? -1.f : 1.f;
    c.vbuf[RR_LDG(&m.limit_dofadr[id & 0xffff])] = sg * frc[r];
  }
  for (int k = c.lane; k < c.nca; k += 32) {
    const float *f = frc + c.nla + 4 * k;
    float mu = RR_LDG(&m.pair_mu[RR_LDG(&m.con_pair[c.cact[k]])]);
    g3[3 * k] = f[0] + f[1] + f[2] + f[3];
    g3[3 * k + 1] = mu * f[0] - mu * f[1];
    g3[3 * k + 2] = mu * f[2] - mu * f[3];
  }
  __syncwarp();
  RR_FOR_S {
    int i = c.lane + 32 * s;
    float acc = 0.f;
    if (i < m.nv) {
      acc = c.vbuf[i];
      for (int k = 0; k < c.nca; k++) {
        int cc = c.cact[k];
        int ld = RR_LDG(&m.pair_lastdof[RR_LDG(&m.con_pair[cc])]);
        if (i <= ld && ld <= i + c.nd[s]) {
          int len = RR_LDG(&m.dof_depth[ld]) + 1;
          const float *J = c.con_J + RR_LDG(&m.con_Jadr[cc]) + c.dep[s];
          acc += J[0] * g3[3 * k] + J[len] * g3[3 * k + 1] + J[2 * len] * g3[3 * k + 2];
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
  /* joint limits */
  int nla = 0;
  for (int base = 0; base < m.nlimit; base += 32) {
    int l = base + c.lane;
    bool active = false;
    float pos = 0.f, dlo = 0.f, dhi = 0.f;
    if (l < m.nlimit) {
      float q = c.qpos[RR_LDG(&m.limit_qposadr[l])];
      dlo = q - RR_LDG(&m.limit_range[2 * l]);
      dhi = RR_LDG(&m.limit_range[2 * l + 1]) - q;
      pos = fminf(dlo, dhi) - RR_LDG(&m.limit_margin[l]);
      active = pos < 0.f;
    }
    unsigned mask = __ballot_sync(RR_FULL, active);
    if (active) {
      int r = nla + __popc(mask & lt);
      float sr[2] = {RR_LDG(&m.limit_solref[2 * l]), RR_LDG(&m.limit_solref[2 * l + 1])}, si[5], k, b, imp;
#pragma unroll
      for (int q = 0; q < 5; q++) si[q] = RR_LDG(&m.limit_solimp[5 * l + q]);
      kbi(m.timestep, sr, si, pos, k, b, imp);
      float R = fmaxf(RR_LDG(&m.limit_invweight[l]) * (1.f - imp) / imp, RR_MINVAL);
      c.row_id[r] = l | (dlo < dhi ? 0 : RR_SIGN_BIT);
      c.row_D[r] = 1.f / R;
      c.row_aref[r] = k * imp * pos; /* temp: completed below */
      c.row_Jaref[r] = b;            /* temp */
    }
    nla += __popc(mask);
  }
  /* contacts */
  int nca = 0;
  for (int base = 0; base < m.ncon; base += 32) {
    int cc = base + c.lane;
    bool active = false;
    if (cc < m.ncon) active = (c.con_dist[cc] - RR_LDG(&m.pair_margin[RR_LDG(&m.con_pair[cc])])) < 0.f;
    unsigned mask = __ballot_sync(RR_FULL, active);
    if (active) c.cact[nca + __popc(mask & lt)] = cc;
    nca += __popc(mask);
  }
  c.nla = nla;
  c.nca = nca;
  __syncwarp();
  /* contact Jacobian blocks (3 x chain) and row parameters */
  for (int k = 0; k < nca; k++) {
    int cc = c.cact[k];
    int p = RR_LDG(&m.con_pair[cc]);
    int ld = RR_LDG(&m.pair_lastdof[p]);
    int len = RR_LDG(&m.dof_depth[ld]) + 1, adr = RR_LDG(&m.dof_rowadr[ld]);
    int rs = RR_LDG(&m.body_rootslot[RR_LDG(&m.pair_body[p])]);
    float off[3], fr[9];
#pragma unroll
    for (int q = 0; q < 3; q++) off[q] = c.con_pos[3 * cc + q] - c.com[3 * rs + q];
#pragma unroll
    for (int q = 0; q < 9; q++) fr[q] = c.con_frame[9 * cc + q];
    float *J = c.con_J + RR_LDG(&m.con_Jadr[cc]);
    for (int t = c.lane; t < len; t += 32) {
      int d = RR_LDG(&m.M_colind[adr + t]);
      float cd[6], cr[3], jp[3];
#pragma unroll
      for (int q = 0; q < 6; q++) cd[q] = c.cdof[6 * d + q];
      cross3(cr, cd, off);
#pragma unroll
      for (int q = 0; q < 3; q++) jp[q] = cd[3 + q] + cr[q];
#pragma unroll
      for (int r3 = 0; r3 < 3; r3++) J[r3 * len + t] = fr[3 * r3] * jp[0] + fr[3 * r3 + 1] * jp[1] + fr[3 * r3 + 2] * jp[2];
    }
    if (c.lane < 4) {
      float pos = c.con_dist[cc] - RR_LDG(&m.pair_margin[p]);
      float sr[2] = {RR_LDG(&m.pair_solref[2 * p]), RR_LDG(&m.pair_solref[2 * p + 1])}, si[5], kk, b, imp;
#pragma unroll
      for (int q = 0; q < 5; q++) si[q] = RR_LDG(&m.pair_solimp[5 * p + q]);
      kbi(m.timestep, sr, si, pos, kk, b, imp);
      float mu = RR_LDG(&m.pair_mu[p]), t = RR_LDG(&m.pair_invweight[p]);
      float invw = (t + mu * mu * t) * 2.f * mu * mu / m.impratio;
      float R = fmaxf(invw * (1.f - imp) / imp, RR_MINVAL);
      int r = nla + 4 * k + c.lane;
      c.row_id[r] = m.nlimit + 4 * cc + c.lane;
      c.row_D[r] = 1.f / R;
      c.row_aref[r] = kk * imp * pos;
      c.row_Jaref[r] = b;
    }
  }
  __syncwarp();
  /* debug: dense efc_J / efc_D in MJX row order (inactive rows are zero there) */
  if (c.dbg) {
    float *dJ = c.dbg + dbg_offset(m, RR_DBG_EFC_J), *dD = c.dbg + dbg_offset(m, RR_DBG_EFC_D);
    for (int i = c.lane; i < m.nefc * m.nv; i += 32) dJ[i] = 0.f;
    for (int i = c.lane; i < m.nefc; i += 32) dD[i] = 1.f / RR_MINVAL;
    __syncwarp();
    for (int r = c.lane; r < nla; r += 32) {
      int id = c.row_id[r], l = id & 0xffff;
      dJ[l * m.nv + RR_LDG(&m.limit_dofadr[l])] = (id & RR_SIGN_BIT) ? -1.f : 1.f;
      dD[l] = c.row_D[r];
    }
    for (int k = 0; k < nca; k++) {
      int cc = c.cact[k];
      int p = RR_LDG(&m.con_pair[cc]);
      int ld = RR_LDG(&m.pair_lastdof[p]);
      int len = RR_LDG(&m.dof_depth[ld]) + 1, adr = RR_LDG(&m.dof_rowadr[ld]);
      const float *J = c.con_J + RR_LDG(&m.con_Jadr[cc]);
      float mu = RR_LDG(&m.pair_mu[p]);
      for (int t = c.lane; t < len; t += 32) {
        int d = RR_LDG(&m.M_colind[adr + t]);
        for (int q = 0; q < 4; q++) {
          float f = (q & 1) ? -mu : mu;
          dJ[(m.nlimit + 4 * cc + q) * m.nv + d] = J[t] + J[(1 + (q >> 1)) * len + t] * f;
        }
      }
      if (c.lane < 4) dD[m.nlimit + 4 * cc + c.lane] = c.row_D[nla + 4 * k + c.lane];
    }
    __syncwarp();
  }
  /* aref = -b (J qvel) - k imp pos */
  int nra = nla + 4 * nca;
  for (int i = c.lane; i < m.nv; i += 32) c.vbuf[i] = c.qvel[i];
  __syncwarp();
  mul_j<NS>(c, c.row_jv);
  for (int r = c.lane; r < nra; r += 32) c.row_aref[r] = -c.row_Jaref[r] * c.row_jv[r] - c.row_aref[r];
  __syncwarp();
  if (c.dbg) {
    float *dA = c.dbg + dbg_offset(m, RR_DBG_EFC_AREF);
    for (int i = c.lane; i < m.nefc; i += 32) dA[i] = 0.f;
    __syncwarp();
    for (int r = c.lane; r < nra; r += 32) dA[c.row_id[r] & 0xffff] = c.row_aref[r];
    __syncwarp();
  }
}

/* ------------------------------------------------------------------------------------------ solver (B.7) */
struct LSPoint { float alpha, cost, d0, d1; };

template <int NS>
RR_DEV LSPoint ls_eval(Ctx<NS> &c, int nra, float alpha, float g0, float g1, float g2) {
  float q0 = 0.f, q1 = 0.f, q2 = 0.f;
  for (int r = c.lane; r < nra; r += 32) {
    float ja = c.row_Jaref[r], jv = c.row_jv[r], D = c.row_D[r];
    if (ja + alpha * jv < 0.f) {
      q0 += 0.5f * ja * ja * D; q1 += jv * ja * D; q2 += 0.5f * jv * jv * D;
    }
  }
  q0 = g0 + warp_sum(q0); q1 = g1 + warp_sum(q1); q2 = g2 + warp_sum(q2);
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.d0 = 2.f * alpha * q2 + q1;
  p.d1 = 2.f * q2 + (q2 == 0.f ? RR_MINVAL : 0.f);
  return p;
}

/* Given qacc (regs): Ma = M qacc, Jaref = J qacc - aref (rows, smem). */
template <int NS>
RR_DEV void ctx_init(Ctx<NS> &c, const float (&qacc)[NS], float (&Ma)[NS]) {
  mul_m<NS>(c, Ma, qacc); /* leaves qacc staged in vbuf */
  vstore<NS>(c, qacc, c.vbuf);
  __syncwarp();
  mul_j<NS>(c, c.row_Jaref);
  int nra = c.nla + 4 * c.nca;
  for (int r = c.lane; r < nra; r += 32) c.row_Jaref[r] -= c.row_aref[r];
  __syncwarp();
}

/* _update_constraint: forces (into row_jv), qfrc_constraint, returns total cost; gauss out */
template <int NS>
RR_DEV float update_constraint(Ctx<NS> &c, const float (&qacc)[NS], const float (&Ma)[NS], float (&qfc)[NS], float &gauss,
                               bool need_force) {
  int nra = c.nla + 4 * c.nca;
  float cost = 0.f;
  for (int r = c.lane; r < nra; r += 32) {
    float ja = c.row_Jaref[r], D = c.row_D[r];
    bool act = ja < 0.f;
    c.row_jv[r] = act ? D * -ja : 0.f;
    if (act) cost += D * ja * ja;
  }
  cost = 0.5f * warp_sum(cost);
  float g = 0.f;
  RR_FOR_S g += (Ma[s] - c.qfrc_smooth[s]) * (qacc[s] - c.qacc_smooth[s]);
  gauss = 0.5f * warp_sum(g);
  if (need_force) mul_jt<NS>(c, c.row_jv, qfc);
  return cost + gauss;
}

template <int NS>
RR_DEV void solve_constraints(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  const int nra = c.nla + 4 * c.nca;
  const float nvf = (float)(m.nv > 1 ? m.nv : 1);
  const float scale = 1.f / (m.meaninertia * nvf);
  float Ma[NS], grad[NS], Mgrad[NS], search[NS], mv[NS];
  float gauss, cost, prev_cost;
  /* warm start: keep whichever of qacc_warmstart / qacc_smooth has the lower cost */
  {
    float cs, cw, g;
    ctx_init<NS>(c, c.qacc_smooth, Ma);
    cs = update_constraint<NS>(c, c.qacc_smooth, Ma, c.qfrc_constraint, g, false);
    ctx_init<NS>(c, c.warm, Ma);
    cw = update_constraint<NS>(c, c.warm, Ma, c.qfrc_constraint, g, false);
    if (cw < cs) {
      RR_FOR_S c.qacc[s] = c.warm[s];
    } else {
      RR_FOR_S c.qacc[s] = c.qacc_smooth[s];
      ctx_init<NS>(c, c.qacc, Ma);
    }
  }
  cost = update_constraint<NS>(c, c.qacc, Ma, c.qfrc_constraint, gauss, true);
  prev_cost = INFINITY;
  RR_FOR_S { grad[s] = Ma[s] - c.qfrc_smooth[s] - c.qfrc_constraint[s]; Mgrad[s] = grad[s]; }
  solve_ld<NS>(c, Mgrad);
  RR_FOR_S search[s] = -Mgrad[s];
  prof<NS>(c, RR_PROF_SOLVE_INIT);
  int niter = 0;
  for (;;) {
    if (m.iterations != 1) {
      float improvement = (prev_cost - cost) * scale;
      float gradient = sqrtf(vdot<NS>(grad, grad)) * scale;
      if (niter >= m.iterations || improvement < m.tolerance || gradient < m.tolerance) break;
    } else if (niter >= 1) {
      break;
    }
    /* ---- linesearch ---- */
    float smag = sqrtf(vdot<NS>(search, search)) * m.meaninertia * nvf;
    float gtol = m.tolerance * m.ls_tolerance * smag;
    mul_m<NS>(c, mv, search); /* search stays staged in vbuf */
    vstore<NS>(c, search, c.vbuf);
    __syncwarp();
    mul_j<NS>(c, c.row_jv);
    float g0 = gauss;
    float g1 = vdot<NS>(search, Ma) - vdot<NS>(search, c.qfrc_smooth);
    float g2 = 0.5f * vdot<NS>(search, mv);
    LSPoint p0 = ls_eval<NS>(c, nra, 0.f, g0, g1, g2);
    LSPoint lo = ls_eval<NS>(c, nra, p0.alpha - p0.d0 / p0.d1, g0, g1, g2), hi;
    if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
    bool swap = true;
    int ls_iter = 0;
    for (;;) {
      bool done = ls_iter >= m.ls_iterations;
      done |= !swap;
      done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
      done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
      if (done) break;
      LSPoint lo_next = ls_eval<NS>(c, nra, lo.alpha - lo.d0 / lo.d1, g0, g1, g2);
      LSPoint hi_next = ls_eval<NS>(c, nra, hi.alpha - hi.d0 / hi.d1, g0, g1, g2);
      LSPoint mid = ls_eval<NS>(c, nra, 0.5f * (lo.alpha + hi.alpha), g0, g1, g2);
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
    bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
    float alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
    if (improved) {
      RR_FOR_S { c.qacc[s] += search[s] * alpha; Ma[s] += mv[s] * alpha; }
      for (int r = c.lane; r < nra; r += 32) c.row_Jaref[r] += c.row_jv[r] * alpha;
      __syncwarp();
    }
    prof<NS>(c, RR_PROF_SOLVE_LS);
    /* ---- update ---- */
    float prev_grad[NS], prev_Mgrad[NS];
    RR_FOR_S { prev_grad[s] = grad[s]; prev_Mgrad[s] = Mgrad[s]; }
    prev_cost = cost;
    cost = update_constraint<NS>(c, c.qacc, Ma, c.qfrc_constraint, gauss, true);
    RR_FOR_S { grad[s] = Ma[s] - c.qfrc_smooth[s] - c.qfrc_constraint[s]; Mgrad[s] = grad[s]; }
    solve_ld<NS>(c, Mgrad);
    float num = 0.f;
    RR_FOR_S num += grad[s] * (Mgrad[s] - prev_Mgrad[s]);
    num = warp_sum(num);
    float den = fmaxf(RR_MINVAL, vdot<NS>(prev_grad, prev_Mgrad));
    float beta = fmaxf(0.f, num / den);
    RR_FOR_S search[s] = -Mgrad[s] + beta * search[s];
    niter++;
    prof<NS>(c, RR_PROF_SOLVE_UPD);
  }
  c.niter = niter;
  RR_FOR_S c.warm[s] = c.qacc[s];
  if (c.dbg) {
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

/* ------------------------------------------------------------------------------------------ forward / euler */
template <int NS>
RR_DEV void forward(Ctx<NS> &c) {
  const RRModelDev &m = c.m;
  kinematics<NS>(c);
  prof<NS>(c, RR_PROF_FK);
  com_pos<NS>(c);
  prof<NS>(c, RR_PROF_COM);
  dbg_copy<NS>(c, RR_DBG_XPOS, c.xpos, 3 * m.nbody);
  dbg_copy<NS>(c, RR_DBG_XQUAT, c.xquat, 4 * m.nbody);
  dbg_copy<NS>(c, RR_DBG_COM, c.com, 3 * m.nroot);
  dbg_copy<NS>(c, RR_DBG_CINERT, c.cinert, 10 * m.nbody);
  dbg_copy<NS>(c, RR_DBG_CDOF, c.cdof, 6 * m.nv);
  crb_and_mass_matrix<NS>(c);
  prof<NS>(c, RR_PROF_QM);
  factor<NS>(c, 0.f);
  prof<NS>(c, RR_PROF_FACTOR);
  dbg_copy<NS>(c, RR_DBG_M, c.M, m.nM);
  dbg_copy<NS>(c, RR_DBG_LD, c.LD, m.nM);
  float qfrc_bias[NS];
  com_vel_and_rne<NS>(c, qfrc_bias);
  prof<NS>(c, RR_PROF_RNE);
  dbg_copy<NS>(c, RR_DBG_CVEL, c.cvel, 6 * m.nbody);
  smooth_forces<NS>(c, qfrc_bias);
  prof<NS>(c, RR_PROF_SMOOTH);
  if (m.nefc == 0) {
    RR_FOR_S { c.qacc[s] = c.qacc_smooth[s]; c.qfrc_constraint[s] = 0.f; }
    return;
  }
  collision<NS>(c);
  prof<NS>(c, RR_PROF_COLLIDE);
  make_constraint<NS>(c);
  prof<NS>(c, RR_PROF_CONSTRAINT);
  solve_constraints<NS>(c);
}

template <int NS>
RR_DEV void euler(Ctx<NS> &c, float &time) {
  const RRModelDev &m = c.m;
  const float dt = m.timestep;
  /* implicit joint damping: (M + dt diag(damping)) qacc = qfrc_smooth + qfrc_constraint */
  factor<NS>(c, dt);
  float qa[NS];
  RR_FOR_S qa[s] = c.qfrc_smooth[s] + c.qfrc_constraint[s];
  solve_ld<NS>(c, qa);
  for (int u = c.lane; u < m.nu; u += 32) {
    if (RR_LDG(&m.act_dyntype[u]) != 0) {
      int aa = RR_LDG(&m.act_actadr[u]);
      c.act[aa] += c.actdot[aa] * dt;
    }
  }
  RR_FOR_S { int i = c.lane + 32 * s; if (i < m.nv) c.qvel[i] += qa[s] * dt; }
  __syncwarp();
  for (int j = c.lane; j < m.njnt; j += 32) {
    int qadr = RR_LDG(&m.jnt_qposadr[j]), da = RR_LDG(&m.jnt_dofadr[j]);
    if (RR_LDG(&m.jnt_type[j]) == RR_JNT_FREE) {
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
}

/* ------------------------------------------------------------------------------------------ one environment */
template <int NS>
RR_DEV void env_run(const RRModelDev &m, const RRStepArgs &a, int env, float *sm, int lane) {
  Ctx<NS> c(m, a, env, sm, lane);
  if (a.prof) c.tprev = RR_CLOCK();
  const size_t e = (size_t)env;
  /* ---- load state ---- */
  for (int i = lane; i < m.nq; i += 32) c.qpos[i] = a.in_qpos[e * m.nq + i];
  for (int i = lane; i < m.nv; i += 32) c.qvel[i] = a.in_qvel[e * m.nv + i];
  for (int i = lane; i < m.na; i += 32) { c.act[i] = a.in_act[e * m.na + i]; c.actdot[i] = 0.f; }
  for (int i = lane; i < m.nu; i += 32) c.ctrl[i] = a.action ? a.action[e * m.nu + i] : 0.f;
  RR_FOR_S { int i = lane + 32 * s; c.warm[s] = i < m.nv ? a.in_warm[e * m.nv + i] : 0.f; }
  float time = a.in_time ? a.in_time[e] : 0.f;
  /* Brax AutoResetWrapper: steps are zeroed where the previous step ended an episode */
  float steps = 0.f;
  if (a.wrap && a.mode == RR_MODE_STEP) steps = (a.in_done[e] != 0.f) ? 0.f : a.in_steps[e];
  __syncwarp();
  prof<NS>(c, RR_PROF_LOAD);
  /* ---- physics ---- */
  if (a.mode == RR_MODE_INIT) {
    forward<NS>(c);
  } else {
    for (int sub = 0; sub < a.nsub; sub++) {
      forward<NS>(c);
      euler<NS>(c, time);
      prof<NS>(c, RR_PROF_EULER);
    }
  }
  __syncwarp();
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
    for (int i = lane; i < m.nq; i += 32) a.qpos[e * m.nq + i] = a.first_qpos[e * m.nq + i];
    for (int i = lane; i < m.nv; i += 32) a.qvel[e * m.nv + i] = a.first_qvel[e * m.nv + i];
    for (int i = lane; i < m.na; i += 32) a.act[e * m.na + i] = a.first_act[e * m.na + i];
    for (int i = lane; i < m.nv; i += 32) a.warm[e * m.nv + i] = a.first_warm[e * m.nv + i];
    if (a.time && lane == 0) a.time[e] = a.first_time ? a.first_time[e] : 0.f;
  } else {
    for (int i = lane; i < m.nq; i += 32) a.qpos[e * m.nq + i] = c.qpos[i];
    for (int i = lane; i < m.nv; i += 32) a.qvel[e * m.nv + i] = c.qvel[i];
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
      for (int i = lane; i < obs_dim; i += 32) o[i] = fo[i];
    } else {
      for (int i = lane; i < m.nq; i += 32) o[i] = c.qpos[i];
      o += m.nq;
      for (int i = lane; i < m.nv; i += 32) o[i] = c.qvel[i];
      o += m.nv;
      for (int i = lane; i < 10 * nb1; i += 32) o[i] = c.cinert[10 + i];
      o += 10 * nb1;
      for (int i = lane; i < 6 * nb1; i += 32) o[i] = c.cvel[6 + i];
      o += 6 * nb1;
      for (int i = lane; i < m.nv; i += 32) o[i] = c.qfrc_act[i];
      o += m.nv;
      if (lane < 3) {
        int ti = cf_new + 1;
        ti = ti < 0 ? 0 : (ti >= t.track_len ? t.track_len - 1 : ti);
        float v[3] = {t.track_pos[3 * ti] - c.qpos[0], t.track_pos[3 * ti + 1] - c.qpos[1], t.track_pos[3 * ti + 2] - c.qpos[2]};
        float R[9], xq[4] = {c.xquat[4], c.xquat[5], c.xquat[6], c.xquat[7]};
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
  }
  /* optional raw outputs of the last forward pass */
  if (a.xpos) for (int i = lane; i < 3 * m.nbody; i += 32) a.xpos[e * 3 * m.nbody + i] = c.xpos[i];
  if (a.xquat) for (int i = lane; i < 4 * m.nbody; i += 32) a.xquat[e * 4 * m.nbody + i] = c.xquat[i];
  if (a.subtree_com) for (int i = lane; i < 3 * m.nroot; i += 32) a.subtree_com[e * 3 * m.nroot + i] = c.com[i];
  if (a.qfrc_actuator) for (int i = lane; i < m.nv; i += 32) a.qfrc_actuator[e * m.nv + i] = c.qfrc_act[i];
  if (a.cinert) for (int i = lane; i < 10 * m.nbody; i += 32) a.cinert[e * 10 * m.nbody + i] = c.cinert[i];
  if (a.cvel) for (int i = lane; i < 6 * m.nbody; i += 32) a.cvel[e * 6 * m.nbody + i] = c.cvel[i];
  if (a.contact_dist && m.nefc) for (int i = lane; i < m.ncon; i += 32) a.contact_dist[e * m.ncon + i] = c.con_dist[i];
  if (a.qacc) RR_FOR_S { int i = lane + 32 * s; if (i < m.nv) a.qacc[e * m.nv + i] = c.qacc[s]; }
  prof<NS>(c, RR_PROF_EPILOGUE);
}

}  // namespace rr

#endif /* RR_KERNELS_INL_ */
