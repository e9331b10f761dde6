/* rr_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A single-environment, straight-line restatement of the physics that the reference obtains from
 * mujoco.mjx 3.1.x through brax (call sites: Rodent_Env_Brax.py:60 pipeline ctor, :87 pipeline_init ->
 * mjx.forward, :101 pipeline_step -> n_frames x mjx.step).  mujoco-mjx / brax / jax are un-vendored,
 * un-pinned third-party dependencies that are absent from /root/reference and from this image, so
 * this file restates the published MJX algorithms in dense form exactly as
 * MJX runs them with `opt.jacobian = dense` (Rodent_Env_Brax.py:49): dense qM, dense Cholesky, dense
 * efc_J with all nefc rows materialised (inactive rows zeroed).  The CUDA library uses a different
 * (tree-sparse, compacted) formulation, which is what makes the comparison meaningful.
 *
 * PARITY UNPINNED: the reference ships no tests or golden vectors for this path and cannot be run
 * here; the oracle is pinned only by the notebook-derived known answers listed in SURVEY.md section 4
 * (tests/test_oracle_kat.py) and by internal invariants (tests/test_oracle_invariants.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  Compiled twice: -DRR_REAL=float (fp32, as JAX x32) and -DRR_REAL=double.
 *
 * API I/O is always double; internal arithmetic is RR_REAL.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rr_model_fields.h"

#ifndef RR_REAL
#define RR_REAL double
#endif
typedef RR_REAL real;

#define MJ_MINVAL ((real)1e-15)
#define MJ_MINIMP ((real)0.0001)
#define MJ_MAXIMP ((real)0.9999)

#define SQRT(x) ((real)sqrt((double)(x)))
#define FABS(x) ((real)fabs((double)(x)))
#define SIN(x) ((real)sin((double)(x)))
#define COS(x) ((real)cos((double)(x)))
#define POW(x, y) ((real)pow((double)(x), (double)(y)))

typedef struct {
  /* sizes / options */
  int nq, nv, nu, na, nbody, njnt, ngeom, npair, ncon, nlimit, nefc;
  int solver, iterations, ls_iterations;
  real timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  /* model tables (views into the owned copies below) */
  const int32_t *field_i[RR_NFIELDS];
  const real *field_f[RR_NFIELDS];
  int count[RR_NFIELDS];
  int32_t *idata;
  real *fdata;
  /* ---- state ---- */
  real *qpos, *qvel, *act, *ctrl, *qacc_warmstart;
  real time;
  /* ---- position-dependent ---- */
  real *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis;
  real *geom_xpos, *geom_xmat;
  real *subtree_com, *cinert, *cdof, *crb;
  real *qM, *qLD; /* dense nv x nv, qLD = lower Cholesky factor */
  real *con_dist, *con_pos, *con_frame;
  real *efc_J, *efc_pos, *efc_invweight, *efc_solref, *efc_solimp, *efc_D, *efc_aref;
  real *actuator_length, *actuator_moment;
  /* ---- velocity-dependent ---- */
  real *cvel, *cdof_dot, *qfrc_passive, *qfrc_bias;
  real *actuator_velocity, *actuator_force, *act_dot, *qfrc_actuator;
  real *qfrc_smooth, *qacc_smooth;
  /* ---- solver ---- */
  real *qacc, *qfrc_constraint, *efc_force, *efc_Jaref, *efc_active;
  int solver_niter;
  real solver_cost;
  long ls_total; /* number of linesearch iterations executed in the last solve */
  /* scratch */
  real *Ma, *grad, *Mgrad, *search, *mv, *jv, *quad, *tmp_nv, *tmp_nv2, *tmp_nefc, *H;
  real *euler_qacc;
} rr_oracle;

/* -------------------------------------------------------------------------------------------- math */
static void quat_mul(real *r, const real *a, const real *b) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_to_mat(real *m, const real *q) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
/* rotate vector by quaternion (mjx math.rotate) */
static void rotate(real *r, const real *v, const real *q) {
  real m[9];
  quat_to_mat(m, q);
  real x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  real y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  real z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void cross(real *r, const real *a, const real *b) {
  real x = a[1] * b[2] - a[2] * b[1];
  real y = a[2] * b[0] - a[0] * b[2];
  real z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static real dot3(const real *a, const real *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static real dotn(const real *a, const real *b, int n) {
  real s = 0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
static real normalize(real *v, int n) {
  real s = SQRT(dotn(v, v, n));
  if (s > 0) for (int i = 0; i < n; i++) v[i] /= s;
  return s;
}
static void axis_angle_quat(real *q, const real *axis, real angle) {
  real s = SIN(angle * (real)0.5);
  q[0] = COS(angle * (real)0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* mjx math.inert_mul: cinert (10) x motion (6: ang, lin) -> force (6) */
static void inert_mul(real *r, const real *i, const real *v) {
  real ang[3], vel[3], c1[3], c2[3];
  ang[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2];
  ang[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2];
  ang[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2];
  cross(c1, i + 6, v + 3);
  cross(c2, i + 6, v);
  for (int k = 0; k < 3; k++) { ang[k] += c1[k]; vel[k] = i[9] * v[3 + k] - c2[k]; }
  for (int k = 0; k < 3; k++) { r[k] = ang[k]; r[3 + k] = vel[k]; }
}
/* mjx math.motion_cross(u, v) */
static void motion_cross(real *r, const real *u, const real *v) {
  real a[3], b[3], c[3];
  cross(a, u, v);
  cross(b, u + 3, v);
  cross(c, u, v + 3);
  for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
/* mjx math.motion_cross_force(v, f) */
static void motion_cross_force(real *r, const real *v, const real *f) {
  real a[3], b[3], c[3];
  cross(a, v, f);
  cross(b, v + 3, f + 3);
  cross(c, v, f + 3);
  for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}

#define FI(o, name) ((o)->field_i[RR_FIELD_##name])
#define FF(o, name) ((o)->field_f[RR_FIELD_##name])

/* -------------------------------------------------------------------------------------------- create */
static real *ralloc(int n) { return (real *)calloc((size_t)(n > 0 ? n : 1), sizeof(real)); }

rr_oracle *rro_create(const int32_t *dir, const int32_t *idata, int ni, const double *fdata, int nf) {
  rr_oracle *o = (rr_oracle *)calloc(1, sizeof(rr_oracle));
  o->idata = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ni > 0 ? ni : 1));
  memcpy(o->idata, idata, sizeof(int32_t) * (size_t)ni);
  o->fdata = ralloc(nf);
  for (int i = 0; i < nf; i++) o->fdata[i] = (real)fdata[i];
  int k = 0;
#define RR__I(n) o->field_i[k] = o->idata + dir[2 * k]; o->count[k] = dir[2 * k + 1]; k++;
#define RR__F(n) o->field_f[k] = o->fdata + dir[2 * k]; o->count[k] = dir[2 * k + 1]; k++;
  RR_MODEL_FIELDS(RR__I, RR__F)
#undef RR__I
#undef RR__F
  const int32_t *oi = FI(o, opt_i);
  const real *of = FF(o, opt_f);
  o->nq = oi[RR_OI_NQ]; o->nv = oi[RR_OI_NV]; o->nu = oi[RR_OI_NU]; o->na = oi[RR_OI_NA];
  o->nbody = oi[RR_OI_NBODY]; o->njnt = oi[RR_OI_NJNT]; o->ngeom = oi[RR_OI_NGEOM];
  o->npair = oi[RR_OI_NPAIR]; o->ncon = oi[RR_OI_NCON]; o->nlimit = oi[RR_OI_NLIMIT]; o->nefc = oi[RR_OI_NEFC];
  o->solver = oi[RR_OI_SOLVER]; o->iterations = oi[RR_OI_ITERATIONS]; o->ls_iterations = oi[RR_OI_LS_ITERATIONS];
  o->timestep = of[RR_OF_TIMESTEP];
  o->gravity[0] = of[RR_OF_GRAVITY_X]; o->gravity[1] = of[RR_OF_GRAVITY_Y]; o->gravity[2] = of[RR_OF_GRAVITY_Z];
  o->tolerance = of[RR_OF_TOLERANCE]; o->ls_tolerance = of[RR_OF_LS_TOLERANCE];
  o->impratio = of[RR_OF_IMPRATIO]; o->meaninertia = of[RR_OF_MEANINERTIA];
  int nq = o->nq, nv = o->nv, nu = o->nu, nb = o->nbody, nj = o->njnt, ng = o->ngeom, nc = o->ncon, ne = o->nefc;
  o->qpos = ralloc(nq); o->qvel = ralloc(nv); o->act = ralloc(o->na); o->ctrl = ralloc(nu);
  o->qacc_warmstart = ralloc(nv);
  o->xpos = ralloc(3 * nb); o->xquat = ralloc(4 * nb); o->xmat = ralloc(9 * nb); o->xipos = ralloc(3 * nb);
  o->ximat = ralloc(9 * nb); o->xanchor = ralloc(3 * nj); o->xaxis = ralloc(3 * nj);
  o->geom_xpos = ralloc(3 * ng); o->geom_xmat = ralloc(9 * ng);
  o->subtree_com = ralloc(3 * nb); o->cinert = ralloc(10 * nb); o->cdof = ralloc(6 * nv); o->crb = ralloc(10 * nb);
  o->qM = ralloc(nv * nv); o->qLD = ralloc(nv * nv);
  o->con_dist = ralloc(nc); o->con_pos = ralloc(3 * nc); o->con_frame = ralloc(9 * nc);
  o->efc_J = ralloc(ne * nv); o->efc_pos = ralloc(ne); o->efc_invweight = ralloc(ne);
  o->efc_solref = ralloc(2 * ne); o->efc_solimp = ralloc(5 * ne); o->efc_D = ralloc(ne); o->efc_aref = ralloc(ne);
  o->actuator_length = ralloc(nu); o->actuator_moment = ralloc(nu * nv);
  o->cvel = ralloc(6 * nb); o->cdof_dot = ralloc(6 * nv); o->qfrc_passive = ralloc(nv); o->qfrc_bias = ralloc(nv);
  o->actuator_velocity = ralloc(nu); o->actuator_force = ralloc(nu); o->act_dot = ralloc(o->na);
  o->qfrc_actuator = ralloc(nv); o->qfrc_smooth = ralloc(nv); o->qacc_smooth = ralloc(nv);
  o->qacc = ralloc(nv); o->qfrc_constraint = ralloc(nv); o->efc_force = ralloc(ne); o->efc_Jaref = ralloc(ne);
  o->efc_active = ralloc(ne);
  o->Ma = ralloc(nv); o->grad = ralloc(nv); o->Mgrad = ralloc(nv); o->search = ralloc(nv); o->mv = ralloc(nv);
  o->jv = ralloc(ne); o->quad = ralloc(3 * ne); o->tmp_nv = ralloc(nv); o->tmp_nv2 = ralloc(nv);
  o->tmp_nefc = ralloc(ne); o->H = ralloc(nv * nv); o->euler_qacc = ralloc(nv);
  /* mjx.make_data: qpos = qpos0, everything else zero */
  for (int i = 0; i < nq; i++) o->qpos[i] = FF(o, qpos0)[i];
  return o;
}

void rro_destroy(rr_oracle *o) {
  if (!o) return;
  /* all buffers are leaked-on-purpose-free: free them */
  real **bufs[] = {&o->qpos, &o->qvel, &o->act, &o->ctrl, &o->qacc_warmstart, &o->xpos, &o->xquat, &o->xmat, &o->xipos,
                   &o->ximat, &o->xanchor, &o->xaxis, &o->geom_xpos, &o->geom_xmat, &o->subtree_com, &o->cinert, &o->cdof,
                   &o->crb, &o->qM, &o->qLD, &o->con_dist, &o->con_pos, &o->con_frame, &o->efc_J, &o->efc_pos,
                   &o->efc_invweight, &o->efc_solref, &o->efc_solimp, &o->efc_D, &o->efc_aref, &o->actuator_length,
                   &o->actuator_moment, &o->cvel, &o->cdof_dot, &o->qfrc_passive, &o->qfrc_bias, &o->actuator_velocity,
                   &o->actuator_force, &o->act_dot, &o->qfrc_actuator, &o->qfrc_smooth, &o->qacc_smooth, &o->qacc,
                   &o->qfrc_constraint, &o->efc_force, &o->efc_Jaref, &o->efc_active, &o->Ma, &o->grad, &o->Mgrad,
                   &o->search, &o->mv, &o->jv, &o->quad, &o->tmp_nv, &o->tmp_nv2, &o->tmp_nefc, &o->H, &o->euler_qacc};
  for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); i++) free(*bufs[i]);
  free(o->idata);
  free(o->fdata);
  free(o);
}

void rro_set_options(rr_oracle *o, int solver, int iterations, int ls_iterations) {
  /* Rodent_Env_Brax.py:42-47 */
  o->solver = solver; o->iterations = iterations; o->ls_iterations = ls_iterations;
}

/* -------------------------------------------------------------------------------------------- smooth.kinematics (Appendix B.1) */
static void kinematics(rr_oracle *o) {
  const int32_t *parent = FI(o, body_parentid), *jntadr = FI(o, body_jntadr), *jntnum = FI(o, body_jntnum);
  const int32_t *jtype = FI(o, jnt_type), *jqadr = FI(o, jnt_qposadr);
  const real *bpos = FF(o, body_pos), *bquat = FF(o, body_quat), *jpos = FF(o, jnt_pos), *jaxis = FF(o, jnt_axis);
  const real *qpos0 = FF(o, qpos0);
  o->xpos[0] = o->xpos[1] = o->xpos[2] = 0;
  o->xquat[0] = 1; o->xquat[1] = o->xquat[2] = o->xquat[3] = 0;
  quat_to_mat(o->xmat, o->xquat);
  for (int b = 1; b < o->nbody; b++) {
    int p = parent[b];
    real pos[3], quat[4], r[3];
    rotate(r, bpos + 3 * b, o->xquat + 4 * p);
    for (int k = 0; k < 3; k++) pos[k] = o->xpos[3 * p + k] + r[k];
    quat_mul(quat, o->xquat + 4 * p, bquat + 4 * b);
    for (int j = jntadr[b]; j < jntadr[b] + jntnum[b]; j++) {
      int a = jqadr[j];
      real *anchor = o->xanchor + 3 * j, *axis = o->xaxis + 3 * j;
      if (jtype[j] == RR_JNT_FREE) {
        for (int k = 0; k < 3; k++) anchor[k] = o->qpos[a + k];
        axis[0] = 0; axis[1] = 0; axis[2] = 1;
        for (int k = 0; k < 3; k++) pos[k] = o->qpos[a + k];
        for (int k = 0; k < 4; k++) quat[k] = o->qpos[a + 3 + k];
        normalize(quat, 4);
        for (int k = 0; k < 4; k++) o->qpos[a + 3 + k] = quat[k]; /* normalised quaternion is written back */
      } else {                                                      /* hinge */
        rotate(r, jpos + 3 * j, quat);
        for (int k = 0; k < 3; k++) anchor[k] = r[k] + pos[k];
        rotate(axis, jaxis + 3 * j, quat);
        real qloc[4], q2[4];
        axis_angle_quat(qloc, jaxis + 3 * j, o->qpos[a] - qpos0[a]);
        quat_mul(q2, quat, qloc);
        memcpy(quat, q2, sizeof(q2));
        rotate(r, jpos + 3 * j, quat);
        for (int k = 0; k < 3; k++) pos[k] = anchor[k] - r[k];
      }
    }
    memcpy(o->xpos + 3 * b, pos, sizeof(pos));
    memcpy(o->xquat + 4 * b, quat, sizeof(quat));
    quat_to_mat(o->xmat + 9 * b, quat);
  }
  const real *ipos = FF(o, body_ipos), *iquat = FF(o, body_iquat);
  for (int b = 0; b < o->nbody; b++) {
    real r[3], q[4];
    rotate(r, ipos + 3 * b, o->xquat + 4 * b);
    for (int k = 0; k < 3; k++) o->xipos[3 * b + k] = o->xpos[3 * b + k] + r[k];
    quat_mul(q, o->xquat + 4 * b, iquat + 4 * b);
    quat_to_mat(o->ximat + 9 * b, q);
  }
  const int32_t *gbody = FI(o, geom_bodyid);
  const real *gpos = FF(o, geom_pos), *gquat = FF(o, geom_quat);
  for (int g = 0; g < o->ngeom; g++) {
    int b = gbody[g];
    real r[3], q[4];
    rotate(r, gpos + 3 * g, o->xquat + 4 * b);
    for (int k = 0; k < 3; k++) o->geom_xpos[3 * g + k] = o->xpos[3 * b + k] + r[k];
    quat_mul(q, o->xquat + 4 * b, gquat + 4 * g);
    quat_to_mat(o->geom_xmat + 9 * g, q);
  }
}

/* -------------------------------------------------------------------------------------------- smooth.com_pos (B.2) */
static void com_pos(rr_oracle *o) {
  int nb = o->nbody;
  const int32_t *parent = FI(o, body_parentid), *rootid = FI(o, body_rootid);
  const real *mass = FF(o, body_mass), *inertia = FF(o, body_inertia);
  real *pos = ralloc(3 * nb), *m = ralloc(nb);
  for (int b = 0; b < nb; b++) {
    for (int k = 0; k < 3; k++) pos[3 * b + k] = o->xipos[3 * b + k] * mass[b];
    m[b] = mass[b];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = parent[b];
    for (int k = 0; k < 3; k++) pos[3 * p + k] += pos[3 * b + k];
    m[p] += m[b];
  }
  for (int b = 0; b < nb; b++)
    for (int k = 0; k < 3; k++)
      o->subtree_com[3 * b + k] = (m[b] < MJ_MINVAL) ? o->xipos[3 * b + k] : pos[3 * b + k] / m[b];
  free(pos);
  free(m);
  for (int b = 0; b < nb; b++) {
    const real *R = o->ximat + 9 * b, *I = inertia + 3 * b;
    const real *rc = o->subtree_com + 3 * rootid[b];
    real off[3], *ci = o->cinert + 10 * b, mb = mass[b];
    for (int k = 0; k < 3; k++) off[k] = o->xipos[3 * b + k] - rc[k];
    /* (R diag(I) R^T) + m (|off|^2 1 - off off^T) */
    real A[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += R[3 * r + k] * I[k] * R[3 * c + k];
        A[3 * r + c] = s;
      }
    real d2 = dot3(off, off);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) A[3 * r + c] += mb * ((r == c ? d2 : 0) - off[r] * off[c]);
    ci[0] = A[0]; ci[1] = A[4]; ci[2] = A[8]; ci[3] = A[1]; ci[4] = A[2]; ci[5] = A[5];
    for (int k = 0; k < 3; k++) ci[6 + k] = off[k] * mb;
    ci[9] = mb;
  }
  /* cdof: [ang; lin] about the subtree COM of the kinematic-tree root */
  const int32_t *jtype = FI(o, jnt_type), *jbody = FI(o, jnt_bodyid), *jdof = FI(o, jnt_dofadr);
  for (int j = 0; j < o->njnt; j++) {
    int b = jbody[j];
    const real *rc = o->subtree_com + 3 * rootid[b];
    real off[3];
    for (int k = 0; k < 3; k++) off[k] = rc[k] - o->xanchor[3 * j + k];
    real *cd = o->cdof + 6 * jdof[j];
    if (jtype[j] == RR_JNT_FREE) {
      for (int d = 0; d < 3; d++)
        for (int k = 0; k < 6; k++) cd[6 * d + k] = (k == 3 + d) ? 1 : 0;
      const real *R = o->xmat + 9 * b;
      for (int d = 0; d < 3; d++) {
        real a[3] = {R[d], R[3 + d], R[6 + d]}; /* column d of xmat = body axis d in world */
        real c[3];
        cross(c, a, off);
        for (int k = 0; k < 3; k++) { cd[6 * (3 + d) + k] = a[k]; cd[6 * (3 + d) + 3 + k] = c[k]; }
      }
    } else {
      real c[3];
      cross(c, o->xaxis + 3 * j, off);
      for (int k = 0; k < 3; k++) { cd[k] = o->xaxis[3 * j + k]; cd[3 + k] = c[k]; }
    }
  }
}

/* -------------------------------------------------------------------------------------------- smooth.crb + factor_m (B.3) */
static int cholesky(real *L, const real *A, int n) {
  /* dense lower Cholesky A = L L^T (jax.scipy.linalg.cho_factor) */
  memcpy(L, A, sizeof(real) * (size_t)(n * n));
  for (int j = 0; j < n; j++) {
    real s = L[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
    if (!(s > 0)) return -1;
    real d = SQRT(s);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      real t = L[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / d;
    }
    for (int i = 0; i < j; i++) L[i * n + j] = 0;
  }
  return 0;
}
static void cho_solve(real *x, const real *L, const real *b, int n) {
  for (int i = 0; i < n; i++) {
    real s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    real s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

static void crb(rr_oracle *o) {
  int nb = o->nbody, nv = o->nv;
  const int32_t *parent = FI(o, body_parentid), *dofbody = FI(o, dof_bodyid), *dofparent = FI(o, dof_parentid);
  const real *arm = FF(o, dof_armature);
  memcpy(o->crb, o->cinert, sizeof(real) * (size_t)(10 * nb));
  for (int b = nb - 1; b > 0; b--)
    for (int k = 0; k < 10; k++) o->crb[10 * parent[b] + k] += o->crb[10 * b + k];
  for (int k = 0; k < 10; k++) o->crb[k] = 0;
  memset(o->qM, 0, sizeof(real) * (size_t)(nv * nv));
  for (int i = 0; i < nv; i++) {
    real f[6];
    inert_mul(f, o->crb + 10 * dofbody[i], o->cdof + 6 * i);
    for (int j = i; j >= 0; j = dofparent[j]) {
      real v = dotn(o->cdof + 6 * j, f, 6);
      o->qM[i * nv + j] = v;
      o->qM[j * nv + i] = v;
    }
    o->qM[i * nv + i] += arm[i];
  }
}

static void mul_m(const rr_oracle *o, real *r, const real *v) {
  int nv = o->nv;
  for (int i = 0; i < nv; i++) r[i] = dotn(o->qM + i * nv, v, nv);
}

/* -------------------------------------------------------------------------------------------- collision_driver (B.4) */
static void plane_sphere(const real *n, const real *ppos, const real *c, real r, real *dist, real *pos) {
  real d[3] = {c[0] - ppos[0], c[1] - ppos[1], c[2] - ppos[2]};
  *dist = dot3(d, n) - r;
  for (int k = 0; k < 3; k++) pos[k] = c[k] - n[k] * (r + (real)0.5 * *dist);
}
static void make_frame(real *frame, const real *a_in) {
  real a[3] = {a_in[0], a_in[1], a_in[2]};
  normalize(a, 3);
  real b[3] = {0, 0, 0};
  if (-0.5 < a[1] && a[1] < 0.5) b[1] = 1; else b[2] = 1;
  real ab = dot3(a, b);
  for (int k = 0; k < 3; k++) b[k] -= a[k] * ab;
  normalize(b, 3);
  for (int k = 0; k < 3; k++) { frame[k] = a[k]; frame[3 + k] = b[k]; }
  cross(frame + 6, a, b);
}

static void collision(rr_oracle *o) {
  const int32_t *fn = FI(o, pair_fn), *g1 = FI(o, pair_geom1), *g2 = FI(o, pair_geom2), *adr = FI(o, pair_conadr);
  const real *gsize = FF(o, geom_size);
  for (int p = 0; p < o->npair; p++) {
    const real *pm = o->geom_xmat + 9 * g1[p], *pp = o->geom_xpos + 3 * g1[p];
    real n[3] = {pm[2], pm[5], pm[8]};
    int c = adr[p], g = g2[p];
    const real *gp = o->geom_xpos + 3 * g, *gm = o->geom_xmat + 9 * g;
    if (fn[p] == RR_PAIR_PLANE_SPHERE) {
      plane_sphere(n, pp, gp, gsize[3 * g], o->con_dist + c, o->con_pos + 3 * c);
      make_frame(o->con_frame + 9 * c, n);
    } else if (fn[p] == RR_PAIR_PLANE_CAPSULE) {
      real axis[3] = {gm[2], gm[5], gm[8]};
      real na = dot3(n, axis), b[3];
      for (int k = 0; k < 3; k++) b[k] = axis[k] - n[k] * na;
      real bn = normalize(b, 3);
      if (bn < 0.5) {
        b[0] = 0;
        if (-0.5 < n[1] && n[1] < 0.5) { b[1] = 1; b[2] = 0; } else { b[1] = 0; b[2] = 1; }
      }
      real frame[9];
      for (int k = 0; k < 3; k++) { frame[k] = n[k]; frame[3 + k] = b[k]; }
      cross(frame + 6, n, b);
      for (int s = 0; s < 2; s++) {
        real cpos[3];
        real sg = s == 0 ? (real)1 : (real)-1;
        for (int k = 0; k < 3; k++) cpos[k] = gp[k] + sg * axis[k] * gsize[3 * g + 1];
        plane_sphere(n, pp, cpos, gsize[3 * g], o->con_dist + c + s, o->con_pos + 3 * (c + s));
        memcpy(o->con_frame + 9 * (c + s), frame, sizeof(frame));
      }
    } else {
      /* plane-ellipsoid (rodent_new / optimized / pair): mjx plane_ellipsoid */
      real size[3] = {gsize[3 * g], gsize[3 * g + 1], gsize[3 * g + 2]};
      real nl[3], sv[3]; /* normal in ellipsoid frame, scaled */
      for (int k = 0; k < 3; k++) nl[k] = gm[k] * n[0] + gm[3 + k] * n[1] + gm[6 + k] * n[2];
      for (int k = 0; k < 3; k++) sv[k] = nl[k] * size[k];
      real nrm = SQRT(dot3(sv, sv));
      real lp[3], wp[3];
      for (int k = 0; k < 3; k++) lp[k] = -(sv[k] / nrm) * size[k]; /* support point in direction -n */
      for (int k = 0; k < 3; k++) wp[k] = gp[k] + gm[3 * k] * lp[0] + gm[3 * k + 1] * lp[1] + gm[3 * k + 2] * lp[2];
      real d[3] = {wp[0] - pp[0], wp[1] - pp[1], wp[2] - pp[2]};
      real dist = dot3(d, n);
      o->con_dist[c] = dist;
      for (int k = 0; k < 3; k++) o->con_pos[3 * c + k] = wp[k] - n[k] * dist * (real)0.5;
      make_frame(o->con_frame + 9 * c, n);
    }
  }
}

/* translational Jacobian of a world point attached to `body` (mjx support.jac): row-major 3 x nv */
static void jac_point(const rr_oracle *o, real *jacp, int body, const real *point) {
  int nv = o->nv;
  const int32_t *lastdof = FI(o, body_lastdof), *dofparent = FI(o, dof_parentid), *rootid = FI(o, body_rootid);
  memset(jacp, 0, sizeof(real) * (size_t)(3 * nv));
  const real *rc = o->subtree_com + 3 * rootid[body];
  real off[3] = {point[0] - rc[0], point[1] - rc[1], point[2] - rc[2]};
  for (int d = lastdof[body]; d >= 0; d = dofparent[d]) {
    const real *cd = o->cdof + 6 * d;
    real c[3];
    cross(c, cd, off);
    for (int k = 0; k < 3; k++) jacp[k * nv + d] = cd[3 + k] + c[k];
  }
}

/* -------------------------------------------------------------------------------------------- constraint.make_constraint (B.5) */
static void kbi(const rr_oracle *o, const real *solref, const real *solimp, real pos, real *k, real *b, real *imp) {
  real timeconst = solref[0], dampratio = solref[1];
  if (timeconst < 2 * o->timestep) timeconst = 2 * o->timestep; /* refsafe */
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = dmin < MJ_MINIMP ? MJ_MINIMP : (dmin > MJ_MAXIMP ? MJ_MAXIMP : dmin);
  dmax = dmax < MJ_MINIMP ? MJ_MINIMP : (dmax > MJ_MAXIMP ? MJ_MAXIMP : dmax);
  if (width < MJ_MINVAL) width = MJ_MINVAL;
  mid = mid < MJ_MINIMP ? MJ_MINIMP : (mid > MJ_MAXIMP ? MJ_MAXIMP : mid);
  if (power < 1) power = 1;
  real kk = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  real bb = 2 / (dmax * timeconst);
  if (solref[0] <= 0) kk = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0) bb = -solref[1] / dmax;
  real x = FABS(pos) / width;
  real ia = (1 / POW(mid, power - 1)) * POW(x, power);
  real ib = 1 - (1 / POW(1 - mid, power - 1)) * POW(1 - x, power);
  real y = x < mid ? ia : ib;
  real im = dmin + y * (dmax - dmin);
  im = im < dmin ? dmin : (im > dmax ? dmax : im);
  if (x > 1) im = dmax;
  *k = kk; *b = bb; *imp = im;
}

static void make_constraint(rr_oracle *o) {
  int nv = o->nv, ne = o->nefc, row = 0;
  memset(o->efc_J, 0, sizeof(real) * (size_t)(ne * nv));
  /* joint limits, in joint order */
  const int32_t *ljnt = FI(o, limit_jntid), *jqadr = FI(o, jnt_qposadr), *jdof = FI(o, jnt_dofadr);
  const real *range = FF(o, jnt_range), *jmargin = FF(o, jnt_margin), *jsolref = FF(o, jnt_solref),
             *jsolimp = FF(o, jnt_solimp), *dinvw = FF(o, dof_invweight0);
  for (int l = 0; l < o->nlimit; l++, row++) {
    int j = ljnt[l];
    real q = o->qpos[jqadr[j]];
    real dmin = q - range[2 * j], dmax = range[2 * j + 1] - q;
    real pos = (dmin < dmax ? dmin : dmax) - jmargin[j];
    real active = pos < 0 ? (real)1 : (real)0;
    o->efc_J[row * nv + jdof[j]] = (dmin < dmax ? (real)1 : (real)-1) * active;
    o->efc_pos[row] = pos * active;
    o->efc_invweight[row] = dinvw[jdof[j]] * active;
    for (int k = 0; k < 2; k++) o->efc_solref[2 * row + k] = jsolref[2 * j + k] * active;
    for (int k = 0; k < 5; k++) o->efc_solimp[5 * row + k] = jsolimp[5 * j + k] * active;
  }
  /* pyramidal contacts, 4 rows each */
  const int32_t *g1 = FI(o, pair_geom1), *g2 = FI(o, pair_geom2), *adr = FI(o, pair_conadr), *gbody = FI(o, geom_bodyid);
  const real *mu = FF(o, pair_friction), *psolref = FF(o, pair_solref), *psolimp = FF(o, pair_solimp),
             *pmargin = FF(o, pair_includemargin), *binvw = FF(o, body_invweight0);
  real *j1 = ralloc(3 * nv), *j2 = ralloc(3 * nv), *dc = ralloc(3 * nv);
  for (int p = 0; p < o->npair; p++) {
    int cend = (p + 1 < o->npair) ? adr[p + 1] : o->ncon;
    int b1 = gbody[g1[p]], b2 = gbody[g2[p]];
    for (int c = adr[p]; c < cend; c++) {
      real dist = o->con_dist[c] - pmargin[p];
      real active = dist < 0 ? (real)1 : (real)0;
      jac_point(o, j1, b1, o->con_pos + 3 * c);
      jac_point(o, j2, b2, o->con_pos + 3 * c);
      const real *fr = o->con_frame + 9 * c;
      for (int r = 0; r < 3; r++)
        for (int d = 0; d < nv; d++) {
          real s = 0;
          for (int k = 0; k < 3; k++) s += fr[3 * r + k] * (j2[k * nv + d] - j1[k * nv + d]);
          dc[r * nv + d] = s;
        }
      real t = binvw[2 * b1] + binvw[2 * b2];
      for (int tdir = 0; tdir < 2; tdir++)
        for (int s = 0; s < 2; s++, row++) {
          real f = s == 0 ? mu[p] : -mu[p];
          for (int d = 0; d < nv; d++) o->efc_J[row * nv + d] = (dc[d] + dc[(1 + tdir) * nv + d] * f) * active;
          o->efc_pos[row] = dist * active;
          o->efc_invweight[row] = ((t + f * f * t) * 2 * f * f / o->impratio) * active;
          for (int k = 0; k < 2; k++) o->efc_solref[2 * row + k] = psolref[2 * p + k] * active;
          for (int k = 0; k < 5; k++) o->efc_solimp[5 * row + k] = psolimp[5 * p + k] * active;
        }
    }
  }
  free(j1); free(j2); free(dc);
  for (int r = 0; r < ne; r++) {
    real k, b, imp;
    kbi(o, o->efc_solref + 2 * r, o->efc_solimp + 5 * r, o->efc_pos[r], &k, &b, &imp);
    real R = o->efc_invweight[r] * (1 - imp) / imp;
    if (R < MJ_MINVAL) R = MJ_MINVAL;
    o->efc_D[r] = 1 / R;
    o->efc_aref[r] = -b * dotn(o->efc_J + r * nv, o->qvel, nv) - k * imp * o->efc_pos[r];
  }
}

/* -------------------------------------------------------------------------------------------- transmission / velocity / forces (B.6) */
static void transmission(rr_oracle *o) {
  const int32_t *ajnt = FI(o, actuator_jntid), *jqadr = FI(o, jnt_qposadr), *jdof = FI(o, jnt_dofadr);
  const real *gear = FF(o, actuator_gear);
  memset(o->actuator_moment, 0, sizeof(real) * (size_t)(o->nu * o->nv));
  for (int u = 0; u < o->nu; u++) {
    int j = ajnt[u];
    o->actuator_length[u] = gear[u] * o->qpos[jqadr[j]];
    o->actuator_moment[u * o->nv + jdof[j]] = gear[u];
  }
}

static void com_vel(rr_oracle *o) {
  const int32_t *parent = FI(o, body_parentid), *jntadr = FI(o, body_jntadr), *jntnum = FI(o, body_jntnum);
  const int32_t *jtype = FI(o, jnt_type), *jdof = FI(o, jnt_dofadr);
  for (int k = 0; k < 6; k++) o->cvel[k] = 0;
  for (int b = 1; b < o->nbody; b++) {
    real cv[6];
    memcpy(cv, o->cvel + 6 * parent[b], sizeof(cv));
    for (int j = jntadr[b]; j < jntadr[b] + jntnum[b]; j++) {
      int d0 = jdof[j];
      if (jtype[j] == RR_JNT_FREE) {
        for (int d = 0; d < 3; d++)
          for (int k = 0; k < 6; k++) cv[k] += o->cdof[6 * (d0 + d) + k] * o->qvel[d0 + d];
        for (int d = 0; d < 3; d++) {
          for (int k = 0; k < 6; k++) o->cdof_dot[6 * (d0 + d) + k] = 0;
          motion_cross(o->cdof_dot + 6 * (d0 + 3 + d), cv, o->cdof + 6 * (d0 + 3 + d));
        }
        for (int d = 3; d < 6; d++)
          for (int k = 0; k < 6; k++) cv[k] += o->cdof[6 * (d0 + d) + k] * o->qvel[d0 + d];
      } else {
        motion_cross(o->cdof_dot + 6 * d0, cv, o->cdof + 6 * d0);
        for (int k = 0; k < 6; k++) cv[k] += o->cdof[6 * d0 + k] * o->qvel[d0];
      }
    }
    memcpy(o->cvel + 6 * b, cv, sizeof(cv));
  }
}

static void passive(rr_oracle *o) {
  const int32_t *jtype = FI(o, jnt_type), *jqadr = FI(o, jnt_qposadr), *jdof = FI(o, jnt_dofadr);
  const real *stiff = FF(o, jnt_stiffness), *qspring = FF(o, qpos_spring), *damp = FF(o, dof_damping);
  for (int i = 0; i < o->nv; i++) o->qfrc_passive[i] = 0;
  for (int j = 0; j < o->njnt; j++) {
    if (jtype[j] == RR_JNT_FREE) {
      /* free-joint springs: stiffness is 0 for <freejoint>; the translational part is kept for completeness */
      for (int k = 0; k < 3; k++)
        o->qfrc_passive[jdof[j] + k] = -stiff[j] * (o->qpos[jqadr[j] + k] - qspring[jqadr[j] + k]);
    } else {
      o->qfrc_passive[jdof[j]] = -stiff[j] * (o->qpos[jqadr[j]] - qspring[jqadr[j]]);
    }
  }
  for (int i = 0; i < o->nv; i++) o->qfrc_passive[i] -= damp[i] * o->qvel[i];
}

static void rne(rr_oracle *o) {
  int nb = o->nbody, nv = o->nv;
  const int32_t *parent = FI(o, body_parentid), *dofadr = FI(o, body_dofadr), *dofnum = FI(o, body_dofnum),
                *dofbody = FI(o, dof_bodyid);
  real *cacc = ralloc(6 * nb), *cfrc = ralloc(6 * nb);
  for (int k = 0; k < 3; k++) { cacc[k] = 0; cacc[3 + k] = -o->gravity[k]; }
  for (int b = 1; b < nb; b++) {
    memcpy(cacc + 6 * b, cacc + 6 * parent[b], sizeof(real) * 6);
    for (int d = dofadr[b]; d < dofadr[b] + dofnum[b]; d++)
      for (int k = 0; k < 6; k++) cacc[6 * b + k] += o->cdof_dot[6 * d + k] * o->qvel[d];
  }
  for (int b = 0; b < nb; b++) {
    real f1[6], f2[6], f3[6];
    inert_mul(f1, o->cinert + 10 * b, cacc + 6 * b);
    inert_mul(f2, o->cinert + 10 * b, o->cvel + 6 * b);
    motion_cross_force(f3, o->cvel + 6 * b, f2);
    for (int k = 0; k < 6; k++) cfrc[6 * b + k] = f1[k] + f3[k];
  }
  for (int b = nb - 1; b > 0; b--)
    for (int k = 0; k < 6; k++) cfrc[6 * parent[b] + k] += cfrc[6 * b + k];
  for (int i = 0; i < nv; i++) o->qfrc_bias[i] = dotn(o->cdof + 6 * i, cfrc + 6 * dofbody[i], 6);
  free(cacc);
  free(cfrc);
}

static void fwd_actuation(rr_oracle *o) {
  int nu = o->nu, nv = o->nv;
  const int32_t *dyntype = FI(o, actuator_dyntype), *gaintype = FI(o, actuator_gaintype),
                *biastype = FI(o, actuator_biastype), *ctrllim = FI(o, actuator_ctrllimited),
                *frclim = FI(o, actuator_forcelimited), *actadr = FI(o, actuator_actadr);
  const real *dynprm = FF(o, actuator_dynprm), *gainprm = FF(o, actuator_gainprm), *biasprm = FF(o, actuator_biasprm),
             *ctrlrange = FF(o, actuator_ctrlrange), *frcrange = FF(o, actuator_forcerange);
  for (int i = 0; i < nv; i++) o->qfrc_actuator[i] = 0;
  for (int u = 0; u < nu; u++) {
    real ctrl = o->ctrl[u];
    if (ctrllim[u]) {
      if (ctrl < ctrlrange[2 * u]) ctrl = ctrlrange[2 * u];
      if (ctrl > ctrlrange[2 * u + 1]) ctrl = ctrlrange[2 * u + 1];
    }
    o->actuator_velocity[u] = dotn(o->actuator_moment + u * nv, o->qvel, nv);
    real ctrl_act = ctrl;
    if (dyntype[u] == 2) { /* filter */
      real tau = dynprm[u] < MJ_MINVAL ? MJ_MINVAL : dynprm[u];
      o->act_dot[actadr[u]] = (ctrl - o->act[actadr[u]]) / tau;
      ctrl_act = o->act[actadr[u]];
    }
    real len = o->actuator_length[u], vel = o->actuator_velocity[u];
    real gain = gainprm[3 * u];
    if (gaintype[u] == 1) gain += gainprm[3 * u + 1] * len + gainprm[3 * u + 2] * vel;
    real bias = 0;
    if (biastype[u] == 1) bias = biasprm[3 * u] + biasprm[3 * u + 1] * len + biasprm[3 * u + 2] * vel;
    real force = gain * ctrl_act + bias;
    if (frclim[u]) {
      if (force < frcrange[2 * u]) force = frcrange[2 * u];
      if (force > frcrange[2 * u + 1]) force = frcrange[2 * u + 1];
    }
    o->actuator_force[u] = force;
    for (int i = 0; i < nv; i++) o->qfrc_actuator[i] += o->actuator_moment[u * nv + i] * force;
  }
}

static void fwd_acceleration(rr_oracle *o) {
  for (int i = 0; i < o->nv; i++) o->qfrc_smooth[i] = o->qfrc_passive[i] - o->qfrc_bias[i] + o->qfrc_actuator[i];
  cho_solve(o->qacc_smooth, o->qLD, o->qfrc_smooth, o->nv);
}

/* -------------------------------------------------------------------------------------------- solver.solve (B.7) */
typedef struct { real alpha, cost, deriv_0, deriv_1; } ls_point;

static void mul_J(const rr_oracle *o, real *r, const real *v) {
  for (int e = 0; e < o->nefc; e++) r[e] = dotn(o->efc_J + e * o->nv, v, o->nv);
}

/* _update_constraint: forces, qfrc_constraint and cost at (qacc, Ma, Jaref) */
static real update_constraint(rr_oracle *o, const real *qacc, const real *Ma, const real *Jaref, real *efc_force,
                              real *qfrc_constraint, real *active_out, real *gauss_out) {
  int nv = o->nv, ne = o->nefc;
  real cost = 0;
  for (int e = 0; e < ne; e++) {
    real a = Jaref[e] < 0 ? (real)1 : (real)0;
    efc_force[e] = o->efc_D[e] * -Jaref[e] * a;
    cost += o->efc_D[e] * Jaref[e] * Jaref[e] * a;
    if (active_out) active_out[e] = a;
  }
  cost *= (real)0.5;
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int e = 0; e < ne; e++) s += o->efc_J[e * nv + i] * efc_force[e];
    qfrc_constraint[i] = s;
  }
  real gauss = 0;
  for (int i = 0; i < nv; i++) gauss += (Ma[i] - o->qfrc_smooth[i]) * (qacc[i] - o->qacc_smooth[i]);
  gauss *= (real)0.5;
  if (gauss_out) *gauss_out = gauss;
  return cost + gauss;
}

static real cost_at(rr_oracle *o, const real *qacc) {
  /* _Context.create(grad=False).cost */
  mul_m(o, o->tmp_nv, qacc);
  mul_J(o, o->tmp_nefc, qacc);
  for (int e = 0; e < o->nefc; e++) o->tmp_nefc[e] -= o->efc_aref[e];
  real *f = ralloc(o->nefc);
  real c = update_constraint(o, qacc, o->tmp_nv, o->tmp_nefc, f, o->tmp_nv2, NULL, NULL);
  free(f);
  return c;
}

static ls_point ls_eval(const rr_oracle *o, real alpha, const real *quad_gauss) {
  real q0 = quad_gauss[0], q1 = quad_gauss[1], q2 = quad_gauss[2];
  for (int e = 0; e < o->nefc; e++) {
    if (o->efc_Jaref[e] + alpha * o->jv[e] < 0) {
      q0 += o->quad[3 * e]; q1 += o->quad[3 * e + 1]; q2 += o->quad[3 * e + 2];
    }
  }
  ls_point p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.deriv_0 = 2 * alpha * q2 + q1;
  p.deriv_1 = 2 * q2 + (q2 == 0 ? MJ_MINVAL : 0);
  return p;
}

static void update_gradient(rr_oracle *o, const real *active) {
  int nv = o->nv, ne = o->nefc;
  for (int i = 0; i < nv; i++) o->grad[i] = o->Ma[i] - o->qfrc_smooth[i] - o->qfrc_constraint[i];
  if (o->solver == 0) {
    cho_solve(o->Mgrad, o->qLD, o->grad, nv);
  } else { /* Newton: H = M + J^T D_active J */
    for (int i = 0; i < nv; i++)
      for (int j = 0; j < nv; j++) {
        real s = o->qM[i * nv + j];
        for (int e = 0; e < ne; e++)
          if (active[e] != 0) s += o->efc_J[e * nv + i] * o->efc_D[e] * o->efc_J[e * nv + j];
        o->H[i * nv + j] = s;
      }
    real *L = ralloc(nv * nv);
    cholesky(L, o->H, nv);
    cho_solve(o->Mgrad, L, o->grad, nv);
    free(L);
  }
}

static void solve(rr_oracle *o) {
  int nv = o->nv, ne = o->nefc;
  real scale = 1 / (o->meaninertia * (real)(nv > 1 ? nv : 1));
  /* warm start */
  real cw = cost_at(o, o->qacc_warmstart), cs = cost_at(o, o->qacc_smooth);
  memcpy(o->qacc, cw < cs ? o->qacc_warmstart : o->qacc_smooth, sizeof(real) * (size_t)nv);
  /* _Context.create */
  mul_m(o, o->Ma, o->qacc);
  mul_J(o, o->efc_Jaref, o->qacc);
  for (int e = 0; e < ne; e++) o->efc_Jaref[e] -= o->efc_aref[e];
  real gauss, cost, prev_cost;
  cost = update_constraint(o, o->qacc, o->Ma, o->efc_Jaref, o->efc_force, o->qfrc_constraint, o->efc_active, &gauss);
  prev_cost = (real)INFINITY;
  update_gradient(o, o->efc_active);
  for (int i = 0; i < nv; i++) o->search[i] = -o->Mgrad[i];
  int niter = 0;
  o->ls_total = 0;
  real *prev_grad = ralloc(nv), *prev_Mgrad = ralloc(nv);
  for (;;) {
    if (o->iterations != 1) { /* lax.while_loop cond; with iterations == 1 the body runs exactly once */
      real improvement = (prev_cost - cost) * scale;
      real gradient = SQRT(dotn(o->grad, o->grad, nv)) * scale;
      if (niter >= o->iterations || improvement < o->tolerance || gradient < o->tolerance) break;
    } else if (niter >= 1) break;
    /* ---- _linesearch ---- */
    real smag = SQRT(dotn(o->search, o->search, nv)) * o->meaninertia * (real)(nv > 1 ? nv : 1);
    real gtol = o->tolerance * o->ls_tolerance * smag;
    mul_m(o, o->mv, o->search);
    mul_J(o, o->jv, o->search);
    real quad_gauss[3];
    quad_gauss[0] = gauss;
    quad_gauss[1] = dotn(o->search, o->Ma, nv) - dotn(o->search, o->qfrc_smooth, nv);
    quad_gauss[2] = (real)0.5 * dotn(o->search, o->mv, nv);
    for (int e = 0; e < ne; e++) {
      o->quad[3 * e] = (real)0.5 * o->efc_Jaref[e] * o->efc_Jaref[e] * o->efc_D[e];
      o->quad[3 * e + 1] = o->jv[e] * o->efc_Jaref[e] * o->efc_D[e];
      o->quad[3 * e + 2] = (real)0.5 * o->jv[e] * o->jv[e] * o->efc_D[e];
    }
    ls_point p0 = ls_eval(o, 0, quad_gauss);
    ls_point lo = ls_eval(o, p0.alpha - p0.deriv_0 / p0.deriv_1, quad_gauss), hi;
    if (lo.deriv_0 < p0.deriv_0) { hi = p0; } else { hi = lo; lo = p0; }
    int swap = 1, ls_iter = 0;
    for (;;) {
      int done = ls_iter >= o->ls_iterations;
      done |= !swap;
      done |= (lo.deriv_0 < 0) && (lo.deriv_0 > -gtol);
      done |= (hi.deriv_0 > 0) && (hi.deriv_0 < gtol);
      if (done) break;
      ls_point lo_next = ls_eval(o, lo.alpha - lo.deriv_0 / lo.deriv_1, quad_gauss);
      ls_point hi_next = ls_eval(o, hi.alpha - hi.deriv_0 / hi.deriv_1, quad_gauss);
      ls_point mid = ls_eval(o, (real)0.5 * (lo.alpha + hi.alpha), quad_gauss);
      int swap_lo_next = (lo.deriv_0 > 0) || (lo.deriv_0 < lo_next.deriv_0);
      if (swap_lo_next) lo = lo_next;
      int swap_lo_mid = (mid.deriv_0 < 0) && (lo.deriv_0 < mid.deriv_0);
      if (swap_lo_mid) lo = mid;
      int swap_hi_next = (hi.deriv_0 < 0) || (hi.deriv_0 > hi_next.deriv_0);
      if (swap_hi_next) hi = hi_next;
      int swap_hi_mid = (mid.deriv_0 > 0) && (hi.deriv_0 > mid.deriv_0);
      if (swap_hi_mid) hi = mid;
      swap = swap_lo_next | swap_lo_mid | swap_hi_next | swap_hi_mid;
      ls_iter++;
      o->ls_total++;
    }
    int improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
    real alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
    if (improved) {
      for (int i = 0; i < nv; i++) { o->qacc[i] += o->search[i] * alpha; o->Ma[i] += o->mv[i] * alpha; }
      for (int e = 0; e < ne; e++) o->efc_Jaref[e] += o->jv[e] * alpha;
    }
    /* ---- update ---- */
    memcpy(prev_grad, o->grad, sizeof(real) * (size_t)nv);
    memcpy(prev_Mgrad, o->Mgrad, sizeof(real) * (size_t)nv);
    prev_cost = cost;
    cost = update_constraint(o, o->qacc, o->Ma, o->efc_Jaref, o->efc_force, o->qfrc_constraint, o->efc_active, &gauss);
    update_gradient(o, o->efc_active);
    if (o->solver == 1) {
      for (int i = 0; i < nv; i++) o->search[i] = -o->Mgrad[i];
    } else { /* Polak-Ribiere */
      real num = 0;
      for (int i = 0; i < nv; i++) num += o->grad[i] * (o->Mgrad[i] - prev_Mgrad[i]);
      real den = dotn(prev_grad, prev_Mgrad, nv);
      if (den < MJ_MINVAL) den = MJ_MINVAL;
      real beta = num / den;
      if (beta < 0) beta = 0;
      for (int i = 0; i < nv; i++) o->search[i] = -o->Mgrad[i] + beta * o->search[i];
    }
    niter++;
  }
  free(prev_grad);
  free(prev_Mgrad);
  o->solver_niter = niter;
  o->solver_cost = cost;
  memcpy(o->qacc_warmstart, o->qacc, sizeof(real) * (size_t)nv);
}

/* -------------------------------------------------------------------------------------------- forward / euler / step */
void rro_forward(rr_oracle *o) {
  kinematics(o);
  com_pos(o);
  crb(o);
  if (cholesky(o->qLD, o->qM, o->nv) != 0) fprintf(stderr, "rr_oracle: qM is not positive definite\n");
  collision(o);
  make_constraint(o);
  transmission(o);
  com_vel(o);
  passive(o);
  rne(o);
  fwd_actuation(o);
  fwd_acceleration(o);
  if (o->nefc == 0) {
    memcpy(o->qacc, o->qacc_smooth, sizeof(real) * (size_t)o->nv);
    return;
  }
  solve(o);
}

static void euler(rr_oracle *o) {
  int nv = o->nv;
  real dt = o->timestep;
  const real *damp = FF(o, dof_damping);
  /* implicit joint damping (eulerdamp): (M + dt diag(damping)) qacc = qfrc_smooth + qfrc_constraint */
  memcpy(o->H, o->qM, sizeof(real) * (size_t)(nv * nv));
  for (int i = 0; i < nv; i++) o->H[i * nv + i] += dt * damp[i];
  real *L = ralloc(nv * nv);
  cholesky(L, o->H, nv);
  for (int i = 0; i < nv; i++) o->tmp_nv[i] = o->qfrc_smooth[i] + o->qfrc_constraint[i];
  cho_solve(o->euler_qacc, L, o->tmp_nv, nv);
  free(L);
  /* _advance */
  const int32_t *dyntype = FI(o, actuator_dyntype), *actadr = FI(o, actuator_actadr);
  for (int u = 0; u < o->nu; u++)
    if (dyntype[u] != 0) o->act[actadr[u]] += o->act_dot[actadr[u]] * dt;
  for (int i = 0; i < nv; i++) o->qvel[i] += o->euler_qacc[i] * dt;
  const int32_t *jtype = FI(o, jnt_type), *jqadr = FI(o, jnt_qposadr), *jdof = FI(o, jnt_dofadr);
  for (int j = 0; j < o->njnt; j++) {
    int a = jqadr[j], d = jdof[j];
    if (jtype[j] == RR_JNT_FREE) {
      for (int k = 0; k < 3; k++) o->qpos[a + k] += dt * o->qvel[d + k];
      real v[3] = {o->qvel[d + 3], o->qvel[d + 4], o->qvel[d + 5]};
      real nrm = normalize(v, 3); /* math.quat_integrate */
      real qr[4], q2[4];
      axis_angle_quat(qr, v, dt * nrm);
      quat_mul(q2, o->qpos + a + 3, qr);
      normalize(q2, 4);
      memcpy(o->qpos + a + 3, q2, sizeof(q2));
    } else {
      o->qpos[a] += dt * o->qvel[d];
    }
  }
  o->time += dt;
}

void rro_step(rr_oracle *o) {
  rro_forward(o);
  euler(o);
}

/* run n mjx.step substeps with ctrl held constant (brax PipelineEnv.pipeline_step, Rodent_Env_Brax.py:101) */
void rro_step_n(rr_oracle *o, int n) {
  for (int i = 0; i < n; i++) rro_step(o);
}

/* -------------------------------------------------------------------------------------------- accessors */
typedef struct { const char *name; real *ptr; int n; } named;

static int lookup(rr_oracle *o, const char *name, real **ptr) {
  int nq = o->nq, nv = o->nv, nu = o->nu, nb = o->nbody, nj = o->njnt, ng = o->ngeom, nc = o->ncon, ne = o->nefc;
  named tab[] = {
      {"qpos", o->qpos, nq}, {"qvel", o->qvel, nv}, {"act", o->act, o->na}, {"ctrl", o->ctrl, nu},
      {"qacc_warmstart", o->qacc_warmstart, nv}, {"xpos", o->xpos, 3 * nb}, {"xquat", o->xquat, 4 * nb},
      {"xmat", o->xmat, 9 * nb}, {"xipos", o->xipos, 3 * nb}, {"ximat", o->ximat, 9 * nb},
      {"xanchor", o->xanchor, 3 * nj}, {"xaxis", o->xaxis, 3 * nj}, {"geom_xpos", o->geom_xpos, 3 * ng},
      {"geom_xmat", o->geom_xmat, 9 * ng}, {"subtree_com", o->subtree_com, 3 * nb}, {"cinert", o->cinert, 10 * nb},
      {"cdof", o->cdof, 6 * nv}, {"crb", o->crb, 10 * nb}, {"qM", o->qM, nv * nv}, {"qLD", o->qLD, nv * nv},
      {"contact_dist", o->con_dist, nc}, {"contact_pos", o->con_pos, 3 * nc}, {"contact_frame", o->con_frame, 9 * nc},
      {"efc_J", o->efc_J, ne * nv}, {"efc_pos", o->efc_pos, ne}, {"efc_D", o->efc_D, ne}, {"efc_aref", o->efc_aref, ne},
      {"efc_force", o->efc_force, ne}, {"efc_Jaref", o->efc_Jaref, ne}, {"efc_active", o->efc_active, ne},
      {"actuator_length", o->actuator_length, nu}, {"actuator_velocity", o->actuator_velocity, nu},
      {"actuator_force", o->actuator_force, nu}, {"act_dot", o->act_dot, o->na}, {"cvel", o->cvel, 6 * nb},
      {"cdof_dot", o->cdof_dot, 6 * nv}, {"qfrc_passive", o->qfrc_passive, nv}, {"qfrc_bias", o->qfrc_bias, nv},
      {"qfrc_actuator", o->qfrc_actuator, nv}, {"qfrc_smooth", o->qfrc_smooth, nv}, {"qacc_smooth", o->qacc_smooth, nv},
      {"qacc", o->qacc, nv}, {"qfrc_constraint", o->qfrc_constraint, nv}, {"euler_qacc", o->euler_qacc, nv},
  };
  for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); i++)
    if (strcmp(tab[i].name, name) == 0) { *ptr = tab[i].ptr; return tab[i].n; }
  return -1;
}

int rro_get(rr_oracle *o, const char *name, double *out, int cap) {
  real *p;
  int n = lookup(o, name, &p);
  if (n < 0) return -1;
  for (int i = 0; i < n && i < cap; i++) out[i] = (double)p[i];
  return n;
}

int rro_set(rr_oracle *o, const char *name, const double *in, int n_in) {
  real *p;
  int n = lookup(o, name, &p);
  if (n < 0 || n != n_in) return -1;
  for (int i = 0; i < n; i++) p[i] = (real)in[i];
  return n;
}

double rro_scalar(rr_oracle *o, const char *name) {
  if (!strcmp(name, "time")) return (double)o->time;
  if (!strcmp(name, "solver_niter")) return (double)o->solver_niter;
  if (!strcmp(name, "solver_cost")) return (double)o->solver_cost;
  if (!strcmp(name, "ls_total")) return (double)o->ls_total;
  if (!strcmp(name, "meaninertia")) return (double)o->meaninertia;
  if (!strcmp(name, "timestep")) return (double)o->timestep;
  return NAN;
}

void rro_set_time(rr_oracle *o, double t) { o->time = (real)t; }
int rro_real_bytes(void) { return (int)sizeof(real); }
