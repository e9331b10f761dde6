/* rr_device.h -- device-side model tables, shared-memory layout and kernel argument blocks.
 *
 * The flat-model blob (include/rr_model_fields.h, produced by brax_rodent_run_b200/mjcf.py) is re-tabulated
 * by rr_model_build.h into the fp32 / int32 SoA tables below, specialised for what the step kernel needs:
 * tree levels for the forward scans, the tree-sparse mass-matrix layout (row i = ancestors of dof i in
 * ascending order, diagonal last), per-pair collision constants with the plane pose folded in, and the
 * per-contact Jacobian chain offsets.  It replaces the device pytree that the reference gets from
 * brax.io.mjcf.load_model / mjx.put_model (Rodent_Env_Brax.py:51).
 */
#ifndef RR_DEVICE_H_
#define RR_DEVICE_H_

#include <stdint.h>

/* integer tables: name, expression for the element count (informative) */
#define RR_DEV_INT_TABLES(X)                                                                              \
  X(body_eparent) X(body_rootslot) X(body_jntadr) X(body_jntnum) X(body_anc)                                                   \
  X(jnt_type) X(jnt_qposadr) X(jnt_dofadr) X(jnt_bodyid)                                                  \
  X(dof_bodyid) X(dof_depth) X(dof_ndesc) X(dof_rowadr) X(dof_log2w) X(dof_pack) X(dof_cbmask) X(dof_descmask) X(dof_ancmask) X(M_meta)                            \
  X(act_dofadr) X(act_qposadr) X(act_dyntype) X(act_gaintype) X(act_biastype) X(act_ctrllimited)          \
  X(act_forcelimited) X(act_actadr)                                                                       \
  X(pair_fn) X(pair_body) X(pair_conadr) X(pair_lastdof) X(pair_cb) X(cb_lastdof) X(cb_conadr) X(cb_conlist) X(con_pair)                          \
  X(limit_qposadr) X(limit_dofadr)

#define RR_DEV_FLOAT_TABLES(X)                                                                            \
  X(body_epos) X(body_equat) X(body_ipos) X(body_iquat) X(body_inertia) X(body_mass)                        \
  X(jnt_pos) X(jnt_axis) X(jnt_stiffness) X(qpos0) X(qpos_spring)                                         \
  X(dof_armature) X(dof_damping)                                                                          \
  X(act_gear) X(act_dynprm) X(act_gainprm) X(act_biasprm) X(act_ctrlrange) X(act_forcerange)              \
  X(pair_gpos) X(pair_gquat) X(pair_size) X(pair_plane_n) X(pair_plane_p) X(pair_mu) X(pair_solref)       \
  X(pair_solimp) X(pair_margin) X(pair_invweight)                                                         \
  X(limit_range) X(limit_margin) X(limit_solref) X(limit_solimp) X(limit_invweight)

/* per-environment shared-memory layout (offsets in floats from the warp's base).  Three regions:
 *   A persistent state; B mass matrix + factor; C a union recycled by phase:
 *     C1 kinematics .. smooth forces : xpos xquat cinert cdof | qfrc_act cvel cacc cfrc  (crb fcrb overlay cvel..cfrc)
 *     C2 collision .. solver         : xpos xquat cdof stay (collision / Jacobians read them); the contact arrays,
 *                                      per-contact six-vectors and constraint rows (capR rows) overlay
 *                                      cinert / qfrc_act / cvel / cacc / cfrc, which are dead by then (their
 *                                      observation slices are written to HBM before the solver in the last substep).
 */
struct RRSmem {
  int qpos, qvel, act, ctrl, actdot, com, vbuf, xq1; /* A */
  int M, LD;                                               /* B */
  int xpos, xquat, cdof;                                   /* C, live through the Jacobian build */
  int cinert, qfrc_act, cvel, cacc, cfrc, crb, fcrb;       /* C1 */
  int con_dist, cab, cscr, cbv, cact, ckidx, row_D; /* C2: cab = 12 floats per contact (frame, point - COM); cscr = 6 floats per
                                                       ACTIVE contact (capA in shared memory); rows: D aref Jaref jv id, capR each */
  int capR, capA;
  int total; /* floats per environment */
};

struct RRModelDev {
  int nq, nv, nu, na, nbody, njnt, ngeom, nM, npair, ncon, nlimit, nefc, nlevel, nroot, ncb;
  int nround; /* pointer-doubling rounds of the tree scans: 2^nround >= nlevel - 1; body_anc = the 2^k-th effective
                 ancestor (0 = world) of body b as byte (k nbody + b) of the table (four per word, RR_BODY_ANC) */
  int solver, iterations, ls_iterations;
  float timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  RRSmem sm;
  /* The tables live in two contiguous device buffers; the kernel stages both into shared memory once per CTA and
   * indexes them through the element offsets o_<table> below (RI / RF macros in rr_kernels.inl). */
  /* Per-dof layout by value: the struct is a __grid_constant__ kernel parameter, so k*[i] with a warp-uniform i is a
   * constant-bank load -- the solve loops index these once per column without touching the LSU.  Byte offsets (x 4) so that
   * a coefficient address is one add: krow4[i] = 4 rowadr[i], kdep4[i] = 4 depth[i].  Zero beyond nv. */
  int32_t krow4[160], kdep4[160];
  float kdtd[160];   /* timestep x dof_damping: the diagonal term of the Euler matrix M + dt diag(damping) (factor2) */
  uint8_t kpar[160]; /* body_parentid by value: the leaf-to-root accumulations read it from the constant bank (no LSU round trip in their serial chain) */
  const int32_t *ibuf;
  const float *fbuf;
  int ni, nf; /* element counts of ibuf / fbuf (multiples of 4) */
#define RR__X(n) int o_##n;
  RR_DEV_INT_TABLES(RR__X)
  RR_DEV_FLOAT_TABLES(RR__X)
#undef RR__X
};

/* run-task parameters (Rodent_Env_Brax.py:21-35,62-69) */
struct RRTask {
  const float *track_pos; /* [track_len, 3] device */
  int track_len;
  float ctrl_cost_weight, healthy_reward, healthy_z_lo, healthy_z_hi;
  int terminate_when_unhealthy;
};

enum { RR_MODE_STEP = 0, RR_MODE_INIT = 1 };

/* debug dump record layout (floats, per environment): see rr_debug_field() */
struct RRDebug {
  float *buf;  /* [B, stride] or null */
  int stride;
};

struct RRStepArgs {
  int B, nsub, mode;
  const float *action; /* [B, nu] raw policy action; null => zeros */
  float *qpos, *qvel, *act, *warm, *time; /* in/out state, env-major rows */
  int *cur_frame;                         /* [B] */
  const float *in_qpos, *in_qvel, *in_act, *in_warm, *in_time, *in_done, *in_steps; /* null => in place */
  const int *in_cur_frame;
  RRTask task;
  float *obs; /* [B, obs_dim] */
  float *reward, *done, *metrics; /* [B], [B], [B,3] (pos_reward, reward_quadctrl, reward_alive) */
  /* fused Brax training wrappers (EpisodeWrapper + AutoResetWrapper); wrap = 0 disables */
  int wrap, episode_length;
  float *steps, *truncation; /* [B] */
  const float *first_qpos, *first_qvel, *first_act, *first_warm, *first_time, *first_obs;
  /* optional extra outputs of the last forward pass (null = skip) */
  float *xpos, *xquat, *subtree_com, *qfrc_actuator, *cinert, *cvel, *contact_dist, *qacc, *contact_pos, *contact_frame;
  int *niter; /* [B] solver iterations executed in the last substep */
  float *work;          /* [B] clock cycles spent on this environment (load-balancing hint) or null */
  const int *env_order; /* [slots] slot -> env (-1 idle) or null */
  RRDebug dbg;
  float *scratch;     /* [warp slots, scratch_stride] global overflow for contact Jacobians / constraint rows */
  int scratch_stride; /* floats: 5 align4(nefc) + 8 + 6 align4(ncon) + 32 (per-stage cycle sums of the instrumented build) */
  long long *prof; /* [B, RR_NPROF] clock64 deltas or null */
};

enum {
  RR_PROF_LOAD = 0, RR_PROF_FK, RR_PROF_COM, RR_PROF_CRB, RR_PROF_QM, RR_PROF_FACTOR, RR_PROF_VEL, RR_PROF_RNE,
  RR_PROF_SMOOTH, RR_PROF_COLLIDE, RR_PROF_CONSTRAINT, RR_PROF_SOLVE_INIT, RR_PROF_SOLVE_LS, RR_PROF_SOLVE_UPD,
  RR_PROF_EULER, RR_PROF_EPILOGUE,
  RR_PROF_WAIT, RR_PROF_WAIT2, RR_PROF_MULM, RR_PROF_FACTOR2, RR_PROF_LS_PRE, RR_PROF_LS_MULJ, RR_PROF_LS_EVAL2, RR_PROF_LS_LOOP, RR_PROF_COST,
  RR_NPROF
};

#endif /* RR_DEVICE_H_ */
