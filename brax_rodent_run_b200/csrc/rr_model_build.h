/* rr_model_build.h -- host side: flat-model blob -> step-kernel tables (pure C++, no CUDA calls).
 *
 * Input is the (dir, idata, fdata) blob described in include/rr_model_fields.h; output is one int32 and one
 * fp32 buffer holding every table of RRModelDev plus the per-environment shared-memory layout.
 * The uploader (rr_api.cu) copies the two buffers to the device and fixes the pointers up.
 */
#ifndef RR_MODEL_BUILD_H_
#define RR_MODEL_BUILD_H_

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rr_model_fields.h"
#include "rr_device.h"

struct RRHostModel {
  RRModelDev dev;  /* scalars + layout valid; pointers set by rr_host_model_bind() */
  std::vector<int32_t> ibuf;
  std::vector<float> fbuf;
  /* offsets (in elements) of each table, in X-macro order */
  std::vector<int> ioff, foff, icount, fcount;
  int obs_dim;
};

namespace rr_detail {

struct Blob {
  const int32_t *dir, *idata;
  const double *fdata;
  const int32_t *I(int f) const { return idata + dir[2 * f]; }
  const double *F(int f) const { return fdata + dir[2 * f]; }
  int n(int f) const { return dir[2 * f + 1]; }
};

inline void quat_mul(double *r, const double *a, const double *b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
inline void quat_rot(double *r, const double *v, const double *q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double m[9] = {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
                 2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
                 2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
  double a = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double b = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double c = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = a; r[1] = b; r[2] = c;
}

}  // namespace rr_detail

#define RR_FID(n) RR_FIELD_##n

/* Build the tables.  Throws std::runtime_error("NotImplemented: ...") for model features outside the
 * supported subset (mirrors the NotImplementedError that brax.io.mjcf.load_model raises). */
inline void rr_host_model_build(RRHostModel &hm, const int32_t *dir, const int32_t *idata, int ni, const double *fdata,
                                int nf) {
  using namespace rr_detail;
  (void)ni; (void)nf;
  Blob B{dir, idata, fdata};
  RRModelDev &d = hm.dev;
  std::memset(&d, 0, sizeof(d));
  const int32_t *oi = B.I(RR_FID(opt_i));
  const double *of = B.F(RR_FID(opt_f));
  d.nq = oi[RR_OI_NQ]; d.nv = oi[RR_OI_NV]; d.nu = oi[RR_OI_NU]; d.na = oi[RR_OI_NA]; d.nbody = oi[RR_OI_NBODY];
  d.njnt = oi[RR_OI_NJNT]; d.ngeom = oi[RR_OI_NGEOM]; d.nM = oi[RR_OI_NM]; d.npair = oi[RR_OI_NPAIR];
  d.ncon = oi[RR_OI_NCON]; d.nlimit = oi[RR_OI_NLIMIT]; d.nefc = oi[RR_OI_NEFC];
  d.solver = oi[RR_OI_SOLVER]; d.iterations = oi[RR_OI_ITERATIONS]; d.ls_iterations = oi[RR_OI_LS_ITERATIONS];
  d.timestep = (float)of[RR_OF_TIMESTEP];
  d.gravity[0] = (float)of[RR_OF_GRAVITY_X]; d.gravity[1] = (float)of[RR_OF_GRAVITY_Y]; d.gravity[2] = (float)of[RR_OF_GRAVITY_Z];
  d.tolerance = (float)of[RR_OF_TOLERANCE]; d.ls_tolerance = (float)of[RR_OF_LS_TOLERANCE];
  d.impratio = (float)of[RR_OF_IMPRATIO]; d.meaninertia = (float)of[RR_OF_MEANINERTIA];
  const int nq = d.nq, nv = d.nv, nu = d.nu, nb = d.nbody, nj = d.njnt, nM = d.nM, np = d.npair, nc = d.ncon, nl = d.nlimit;
  if (nv > 160) throw std::runtime_error("NotImplemented: nv > 160");
  if (nb > 160) throw std::runtime_error("NotImplemented: nbody > 160");
  if (d.na != nu && d.na != 0) {
    /* mixed stateful / stateless actuators are fine; nothing to check */
  }

  std::vector<std::vector<int32_t>> IT;
  std::vector<std::vector<float>> FT;
#define RR__X(n) std::vector<int32_t> t_##n;
  RR_DEV_INT_TABLES(RR__X)
#undef RR__X
#define RR__X(n) std::vector<float> t_##n;
  RR_DEV_FLOAT_TABLES(RR__X)
#undef RR__X

  auto cpI = [&](std::vector<int32_t> &dst, int f) { dst.assign(B.I(f), B.I(f) + B.n(f)); };
  auto cpF = [&](std::vector<float> &dst, int f) {
    dst.resize(B.n(f));
    for (int i = 0; i < B.n(f); i++) dst[i] = (float)B.F(f)[i];
  };

  /* ---- bodies ---- */
  cpI(t_body_jntadr, RR_FID(body_jntadr));
  cpI(t_body_jntnum, RR_FID(body_jntnum));
  cpF(t_body_ipos, RR_FID(body_ipos));
  cpF(t_body_iquat, RR_FID(body_iquat)); cpF(t_body_inertia, RR_FID(body_inertia)); cpF(t_body_mass, RR_FID(body_mass));
  const int32_t *rootid = B.I(RR_FID(body_rootid)), *parent = B.I(RR_FID(body_parentid)), *lastdof = B.I(RR_FID(body_lastdof));
  /* kinematic-tree roots (world excluded): slot index per body; world gets slot 0 (its cinert is zero anyway) */
  std::vector<int> roots;
  t_body_rootslot.assign(nb, 0);
  for (int b = 1; b < nb; b++) {
    int r = rootid[b];
    auto it = std::find(roots.begin(), roots.end(), r);
    if (it == roots.end()) { roots.push_back(r); it = roots.end() - 1; }
    t_body_rootslot[b] = (int)(it - roots.begin());
  }
  d.nroot = std::max<int>(1, (int)roots.size());
  /* effective parent (nearest ancestor with joints, or the world) and the composed fixed offset to it */
  {
    const double *bpos_d = B.F(RR_FID(body_pos)), *bquat_d = B.F(RR_FID(body_quat));
    const int32_t *bjntnum = B.I(RR_FID(body_jntnum));
    t_body_eparent.assign(nb, 0);
    t_body_epos.assign(3 * nb, 0.f);
    t_body_equat.assign(4 * nb, 0.f);
    t_body_equat[0] = 1.f;
    for (int b = 1; b < nb; b++) {
      double pos[3] = {bpos_d[3 * b], bpos_d[3 * b + 1], bpos_d[3 * b + 2]};
      double quat[4] = {bquat_d[4 * b], bquat_d[4 * b + 1], bquat_d[4 * b + 2], bquat_d[4 * b + 3]};
      int p = parent[b];
      while (p > 0 && bjntnum[p] == 0) { /* fold the welded parent's offset in: T_p o T_b */
        double r[3], q[4];
        quat_rot(r, pos, bquat_d + 4 * p);
        for (int c = 0; c < 3; c++) pos[c] = bpos_d[3 * p + c] + r[c];
        quat_mul(q, bquat_d + 4 * p, quat);
        std::memcpy(quat, q, sizeof(q));
        p = parent[p];
      }
      t_body_eparent[b] = p;
      for (int c = 0; c < 3; c++) t_body_epos[3 * b + c] = (float)pos[c];
      for (int c = 0; c < 4; c++) t_body_equat[4 * b + c] = (float)quat[c];
    }
  }
  /* effective depth of the tree (level 0 = world) */
  std::vector<int> depth(nb, 0);
  int maxd = 0;
  for (int b = 1; b < nb; b++) { depth[b] = depth[t_body_eparent[b]] + 1; maxd = std::max(maxd, depth[b]); }
  d.nlevel = maxd + 1;
  /* pointer-doubling tables for the tree scans (kinematics, com_vel): the 2^k-th effective ancestor of every body.
   * A free-joint body takes its pose from qpos whatever its parent is, so it hangs off the world here. */
  {
    const int32_t *jadr0 = B.I(RR_FID(body_jntadr)), *jnum0 = B.I(RR_FID(body_jntnum)), *jtype0 = B.I(RR_FID(jnt_type));
    for (int b = 1; b < nb; b++)
      for (int j = jadr0[b]; j < jadr0[b] + jnum0[b]; j++)
        if (jtype0[j] == RR_JNT_FREE) {
          if (t_body_eparent[b] != 0) throw std::runtime_error("NotImplemented: free joint below a moving body");
          if (jnum0[b] != 1) throw std::runtime_error("NotImplemented: free joint sharing a body with other joints");
        }
    int R = 0;
    while ((1 << R) < maxd) R++;
    d.nround = R;
    std::vector<int> anc((size_t)std::max(R, 1) * nb, 0);
    for (int b = 0; b < nb; b++) anc[b] = t_body_eparent[b];
    for (int k = 1; k < R; k++)
      for (int b = 0; b < nb; b++) anc[(size_t)k * nb + b] = anc[(size_t)(k - 1) * nb + anc[(size_t)(k - 1) * nb + b]];
    t_body_anc.assign((anc.size() + 3) / 4, 0); /* one byte per entry (nbody <= 160) */
    for (size_t e = 0; e < anc.size(); e++) t_body_anc[e >> 2] |= (int32_t)((uint32_t)anc[e] << (8 * (e & 3)));
  }

  /* ---- joints ---- */
  cpI(t_jnt_type, RR_FID(jnt_type)); cpI(t_jnt_qposadr, RR_FID(jnt_qposadr)); cpI(t_jnt_dofadr, RR_FID(jnt_dofadr));
  cpI(t_jnt_bodyid, RR_FID(jnt_bodyid));
  cpF(t_jnt_pos, RR_FID(jnt_pos)); cpF(t_jnt_axis, RR_FID(jnt_axis)); cpF(t_jnt_stiffness, RR_FID(jnt_stiffness));
  cpF(t_qpos0, RR_FID(qpos0)); cpF(t_qpos_spring, RR_FID(qpos_spring));
  for (int j = 0; j < nj; j++)
    if (t_jnt_type[j] != RR_JNT_FREE && t_jnt_type[j] != RR_JNT_HINGE) throw std::runtime_error("NotImplemented: ball/slide joints");

  /* ---- dofs + tree-sparse layout ---- */
  cpI(t_dof_bodyid, RR_FID(dof_bodyid));
  cpI(t_dof_rowadr, RR_FID(M_rowadr));
  std::vector<int32_t> t_M_colind, t_M_rowid;
  cpI(t_M_colind, RR_FID(M_colind));
  cpF(t_dof_armature, RR_FID(dof_armature)); cpF(t_dof_damping, RR_FID(dof_damping));
  const int32_t *dofparent = B.I(RR_FID(dof_parentid)), *rownnz = B.I(RR_FID(M_rownnz));
  t_dof_depth.assign(nv, 0);
  t_dof_ndesc.assign(nv, 0);
  for (int i = 0; i < nv; i++) {
    t_dof_depth[i] = rownnz[i] - 1;
    for (int k = dofparent[i]; k >= 0; k = dofparent[k]) t_dof_ndesc[k]++;
  }
  /* descendants of a dof must be the contiguous range (i, i + ndesc]: DFS numbering */
  for (int i = 0; i < nv; i++)
    for (int k = i + 1; k < nv; k++) {
      bool anc = false;
      for (int a = dofparent[k]; a >= 0; a = dofparent[a]) if (a == i) anc = true;
      if (anc != (k <= i + t_dof_ndesc[i])) throw std::runtime_error("internal: dof numbering is not depth-first");
    }
  t_M_rowid.assign(nM, 0);
  for (int i = 0; i < nv; i++)
    for (int t = 0; t < rownnz[i]; t++) t_M_rowid[t_dof_rowadr[i] + t] = i;
  if (t_M_rowid.empty()) { t_M_rowid.push_back(0); t_M_colind.push_back(0); }
  /* packed entry metadata (row | col << 8 | rowadr[col] << 16) and the triangular enumeration used by factor() */
  if (nv > 255 || nM > 65535) throw std::runtime_error("NotImplemented: nv > 255");
  t_M_meta.resize(t_M_colind.size());
  for (size_t e = 0; e < t_M_colind.size(); e++)
    t_M_meta[e] = nv ? (t_M_rowid[e] | (t_M_colind[e] << 8) | (t_dof_rowadr[t_M_colind[e]] << 16)) : 0;
  /* packed dof metadata: rowadr | depth << 16 | ndesc << 24 */
  t_dof_pack.resize(nv);
  for (int i = 0; i < nv; i++) t_dof_pack[i] = t_dof_rowadr[i] | (t_dof_depth[i] << 16) | (t_dof_ndesc[i] << 24);
  for (int i = 0; i < 160; i++) { d.krow4[i] = i < nv ? 4 * t_dof_rowadr[i] : 0; d.kdep4[i] = i < nv ? 4 * t_dof_depth[i] : 0; }
  for (int b = 0; b < 160; b++) d.kpar[b] = (uint8_t)(b < nb ? parent[b] : 0);
  for (int i = 0; i < 160; i++) d.kdtd[i] = i < nv ? d.timestep * t_dof_damping[i] : 0.f;
  /* per dof and block of 32 columns: which columns are descendants / ancestors of the dof (the solves' predicates) */
  {
    const int nb32 = (nv + 31) / 32;
    t_dof_descmask.assign((size_t)std::max(nv, 1) * std::max(nb32, 1), 0);
    t_dof_ancmask.assign((size_t)std::max(nv, 1) * std::max(nb32, 1), 0);
    for (int j = 0; j < nv; j++)
      for (int i = j + 1; i <= j + t_dof_ndesc[j]; i++) {
        t_dof_descmask[(size_t)j * nb32 + i / 32] |= (int32_t)(1u << (i & 31));
        t_dof_ancmask[(size_t)i * nb32 + j / 32] |= (int32_t)(1u << (j & 31));
      }
  }
  if (t_dof_pack.empty()) t_dof_pack.push_back(0);
  /* factor(): row width rounded up to a power of two (log2), for the lane-group split of short rows */
  t_dof_log2w.assign(nv, 5);
  {
    std::vector<int> nchild(nv, 0);
    for (int i = 0; i < nv; i++) if (dofparent[i] >= 0) nchild[dofparent[i]]++;
    for (int i = 0; i < nv; i++) {
      int lw = 0;
      while ((1 << lw) < t_dof_depth[i] + 1 && lw < 5) lw++;
      /* bits 8-9: T - 1, the number of rows i-1 .. i-T+1 eliminated in the same step as row i (each the parent of the one
       * below with no other child; the row fits one register per lane).  factor2 sweeps from nv - 1 down and takes groups
       * greedily from the leaf end of a chain, so only the entry of a group's deepest row is read. */
      int T = 1;
      if (t_dof_depth[i] < 32)
        while (T < 4 && i - T >= 0 && dofparent[i - T + 1] == i - T && nchild[i - T] == 1 && t_dof_depth[i - T] >= 0) T++;
      t_dof_log2w[i] = lw | ((T - 1) << 8);
    }
  }
  {
    int maxdep = 0;
    for (int i = 0; i < nv; i++) maxdep = std::max(maxdep, t_dof_depth[i]);
    if (maxdep > 63) throw std::runtime_error("NotImplemented: kinematic chains deeper than 63 dofs");
  }

  /* ---- actuators ---- */
  const int32_t *ajnt = B.I(RR_FID(actuator_jntid));
  t_act_dofadr.resize(nu); t_act_qposadr.resize(nu);
  std::vector<int> used(nv, 0);
  for (int u = 0; u < nu; u++) {
    t_act_dofadr[u] = t_jnt_dofadr[ajnt[u]];
    t_act_qposadr[u] = t_jnt_qposadr[ajnt[u]];
    if (used[t_act_dofadr[u]]++) throw std::runtime_error("NotImplemented: two actuators on one joint");
  }
  cpI(t_act_dyntype, RR_FID(actuator_dyntype)); cpI(t_act_gaintype, RR_FID(actuator_gaintype));
  cpI(t_act_biastype, RR_FID(actuator_biastype)); cpI(t_act_ctrllimited, RR_FID(actuator_ctrllimited));
  cpI(t_act_forcelimited, RR_FID(actuator_forcelimited)); cpI(t_act_actadr, RR_FID(actuator_actadr));
  cpF(t_act_gear, RR_FID(actuator_gear)); cpF(t_act_dynprm, RR_FID(actuator_dynprm));
  cpF(t_act_gainprm, RR_FID(actuator_gainprm)); cpF(t_act_biasprm, RR_FID(actuator_biasprm));
  cpF(t_act_ctrlrange, RR_FID(actuator_ctrlrange)); cpF(t_act_forcerange, RR_FID(actuator_forcerange));

  /* ---- collision pairs: plane (on a static body) vs sphere / capsule / ellipsoid ---- */
  const int32_t *g1 = B.I(RR_FID(pair_geom1)), *g2 = B.I(RR_FID(pair_geom2)), *gbody = B.I(RR_FID(geom_bodyid));
  const double *gpos = B.F(RR_FID(geom_pos)), *gquat = B.F(RR_FID(geom_quat)), *gsize = B.F(RR_FID(geom_size));
  const double *bpos = B.F(RR_FID(body_pos)), *bquat = B.F(RR_FID(body_quat)), *binvw = B.F(RR_FID(body_invweight0));
  cpI(t_pair_fn, RR_FID(pair_fn)); cpI(t_pair_conadr, RR_FID(pair_conadr));
  cpF(t_pair_mu, RR_FID(pair_friction)); cpF(t_pair_solref, RR_FID(pair_solref)); cpF(t_pair_solimp, RR_FID(pair_solimp));
  cpF(t_pair_margin, RR_FID(pair_includemargin));
  t_pair_body.resize(np); t_pair_lastdof.resize(np); t_pair_gpos.resize(3 * np); t_pair_gquat.resize(4 * np);
  t_pair_size.resize(3 * np); t_pair_plane_n.resize(3 * np); t_pair_plane_p.resize(3 * np); t_pair_invweight.resize(np);
  t_con_pair.assign(std::max(nc, 1), 0);
  t_pair_cb.assign(std::max(np, 1), 0);
  std::vector<int> cb_bodies; /* distinct bodies carrying a collision geom ("contact bodies") */
  for (int p = 0; p < np; p++) {
    int bp = gbody[g1[p]], b2 = gbody[g2[p]];
    if (lastdof[bp] >= 0) throw std::runtime_error("NotImplemented: collision plane on a moving body");
    if (lastdof[b2] < 0) throw std::runtime_error("NotImplemented: collision geom on a static body");
    /* world pose of the static plane: compose up the (static) body chain */
    double pq[4] = {1, 0, 0, 0}, pp[3] = {0, 0, 0};
    std::vector<int> chain;
    for (int b = bp; b > 0; b = parent[b]) chain.push_back(b);
    for (int k = (int)chain.size() - 1; k >= 0; k--) {
      int b = chain[k];
      double r[3], q[4];
      quat_rot(r, bpos + 3 * b, pq);
      for (int c = 0; c < 3; c++) pp[c] += r[c];
      quat_mul(q, pq, bquat + 4 * b);
      std::memcpy(pq, q, sizeof(q));
    }
    double r[3], q[4], z[3] = {0, 0, 1}, n[3];
    quat_rot(r, gpos + 3 * g1[p], pq);
    quat_mul(q, pq, gquat + 4 * g1[p]);
    quat_rot(n, z, q);
    for (int c = 0; c < 3; c++) {
      t_pair_plane_p[3 * p + c] = (float)(pp[c] + r[c]);
      t_pair_plane_n[3 * p + c] = (float)n[c];
      t_pair_gpos[3 * p + c] = (float)gpos[3 * g2[p] + c];
      t_pair_size[3 * p + c] = (float)gsize[3 * g2[p] + c];
    }
    for (int c = 0; c < 4; c++) t_pair_gquat[4 * p + c] = (float)gquat[4 * g2[p] + c];
    t_pair_body[p] = b2;
    t_pair_lastdof[p] = lastdof[b2];
    t_pair_invweight[p] = (float)(binvw[2 * bp] + binvw[2 * b2]);
    int cend = (p + 1 < np) ? t_pair_conadr[p + 1] : nc;
    for (int c = t_pair_conadr[p]; c < cend; c++) t_con_pair[c] = p;
    auto it = std::find(cb_bodies.begin(), cb_bodies.end(), b2);
    if (it == cb_bodies.end()) { cb_bodies.push_back(b2); it = cb_bodies.end() - 1; }
    t_pair_cb[p] = (int)(it - cb_bodies.begin());
  }
  d.ncb = (int)cb_bodies.size();
  for (int b : cb_bodies) t_cb_lastdof.push_back(lastdof[b]);
  if (t_cb_lastdof.empty()) t_cb_lastdof.push_back(0);
  /* per dof: bit kb set when the chain of contact body kb contains the dof (mul_jt gathers only over these) */
  if (d.ncb > 64) throw std::runtime_error("NotImplemented: more than 64 bodies with collision geoms");
  if (nc > 255 || np > 4095) throw std::runtime_error("NotImplemented: more than 255 contacts / 4095 collision pairs");
  t_dof_cbmask.assign(2 * std::max(nv, 1), 0);
  for (int kb = 0; kb < d.ncb; kb++)
    for (int dd = t_cb_lastdof[kb]; dd >= 0; dd = dofparent[dd]) t_dof_cbmask[2 * dd + (kb >> 5)] |= (int32_t)(1u << (kb & 31));
  /* static contact list of each contact body */
  t_cb_conadr.assign(d.ncb + 1, 0);
  for (int kb = 0; kb < d.ncb; kb++) {
    t_cb_conadr[kb] = (int)t_cb_conlist.size();
    for (int c = 0; c < nc; c++) if (t_pair_cb[t_con_pair[c]] == kb) t_cb_conlist.push_back(c);
  }
  t_cb_conadr[d.ncb] = (int)t_cb_conlist.size();
  if (t_cb_conlist.empty()) t_cb_conlist.push_back(0);

  /* ---- joint limits ---- */
  const int32_t *ljnt = B.I(RR_FID(limit_jntid));
  const double *jrange = B.F(RR_FID(jnt_range)), *jmargin = B.F(RR_FID(jnt_margin)), *jsolref = B.F(RR_FID(jnt_solref)),
               *jsolimp = B.F(RR_FID(jnt_solimp)), *dinvw = B.F(RR_FID(dof_invweight0));
  t_limit_qposadr.resize(nl); t_limit_dofadr.resize(nl); t_limit_range.resize(2 * nl); t_limit_margin.resize(nl);
  t_limit_solref.resize(2 * nl); t_limit_solimp.resize(5 * nl); t_limit_invweight.resize(nl);
  for (int l = 0; l < nl; l++) {
    int j = ljnt[l];
    t_limit_qposadr[l] = t_jnt_qposadr[j];
    t_limit_dofadr[l] = t_jnt_dofadr[j];
    t_limit_range[2 * l] = (float)jrange[2 * j]; t_limit_range[2 * l + 1] = (float)jrange[2 * j + 1];
    t_limit_margin[l] = (float)jmargin[j];
    for (int k = 0; k < 2; k++) t_limit_solref[2 * l + k] = (float)jsolref[2 * j + k];
    for (int k = 0; k < 5; k++) t_limit_solimp[5 * l + k] = (float)jsolimp[5 * j + k];
    t_limit_invweight[l] = (float)dinvw[t_jnt_dofadr[j]];
  }

  /* ---- pack ---- */
  hm.ibuf.clear(); hm.fbuf.clear(); hm.ioff.clear(); hm.foff.clear(); hm.icount.clear(); hm.fcount.clear();
  auto padI = [&]() { while (hm.ibuf.size() % 4) hm.ibuf.push_back(0); };
  auto padF = [&]() { while (hm.fbuf.size() % 4) hm.fbuf.push_back(0.f); };
#define RR__X(n)                                                   \
  hm.ioff.push_back((int)hm.ibuf.size());                          \
  hm.icount.push_back((int)t_##n.size());                          \
  hm.ibuf.insert(hm.ibuf.end(), t_##n.begin(), t_##n.end());       \
  padI();
  RR_DEV_INT_TABLES(RR__X)
#undef RR__X
#define RR__X(n)                                                   \
  hm.foff.push_back((int)hm.fbuf.size());                          \
  hm.fcount.push_back((int)t_##n.size());                          \
  hm.fbuf.insert(hm.fbuf.end(), t_##n.begin(), t_##n.end());       \
  padF();
  RR_DEV_FLOAT_TABLES(RR__X)
#undef RR__X
  if (hm.ibuf.empty()) hm.ibuf.push_back(0);
  if (hm.fbuf.empty()) hm.fbuf.push_back(0.f);
  padI();
  padF();

  /* ---- shared-memory layout (see RRSmem) ---- */
  RRSmem &s = d.sm;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  s.qpos = take(nq); s.qvel = take(nv); s.act = take(d.na); s.ctrl = take(nu); s.actdot = take(d.na);
  s.com = take(3 * d.nroot); s.vbuf = take(nv);  s.xq1 = take(4);
  s.M = take(nM); s.LD = take(std::max(nM, 8 * nb)); /* LD doubles as the second buffer of the tree scans (8 floats per body) */
  s.xpos = take(3 * nb); s.xquat = take(4 * nb); s.cdof = take(6 * nv);
  const int c0 = o;
  /* C1 */
  s.cinert = take(10 * nb); s.qfrc_act = take(nv);
  const int c1b = o;
  s.cvel = take(6 * nb); s.cacc = take(6 * nb); s.cfrc = s.cacc; /* cfrc is computed in place over cacc */
  int c1_end = o;
  o = c1b; s.crb = take(10 * nb); s.fcrb = s.crb;
  c1_end = std::max(c1_end, o);
  /* factor2 stages 9 floats per descendant of a row in the C1 span (dead by then) */
  if (9 * ((nv + 3) & ~3) > c1_end - c0) throw std::runtime_error("NotImplemented: factorisation staging does not fit (nv >> nbody)");
  /* C2: contact arrays, then as many constraint rows (5 arrays) as fit in the recycled span */
  o = c0;
  s.capA = std::min(nc, 32); /* active contacts whose six-vectors stay in shared memory (more go to the global scratch) */
  s.con_dist = take(nc); s.cab = take(12 * nc); s.cscr = take(6 * s.capA); s.cbv = take(6 * d.ncb); s.cact = take(nc); s.ckidx = take(nc);
  int avail = c1_end - o;
  int capR = std::min((d.nefc + 3) & ~3, std::max(avail / 5, 32) & ~3);
  if (capR < 4) capR = 4;
  s.capR = capR;
  s.row_D = take(5 * capR);
  s.total = std::max(c1_end, o);

  hm.obs_dim = nq + nv + 10 * (nb - 1) + 6 * (nb - 1) + nv + 3; /* Rodent_Env_Brax.py:149-158 */
}

/* Fill the table offsets of `dev` and point it at copies of ibuf / fbuf living at ibase / fbase. */
inline void rr_host_model_bind(const RRHostModel &hm, RRModelDev &dev, const int32_t *ibase, const float *fbase) {
  int k = 0;
#define RR__X(n) dev.o_##n = hm.ioff[k++];
  RR_DEV_INT_TABLES(RR__X)
#undef RR__X
  k = 0;
#define RR__X(n) dev.o_##n = hm.foff[k++];
  RR_DEV_FLOAT_TABLES(RR__X)
#undef RR__X
  dev.ibuf = ibase;
  dev.fbuf = fbase;
  dev.ni = (int)hm.ibuf.size();
  dev.nf = (int)hm.fbuf.size();
}

#endif /* RR_MODEL_BUILD_H_ */
