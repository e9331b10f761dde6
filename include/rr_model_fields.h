/* rr_model_fields.h -- single source of truth for the flat-model blob layout.
 *
 * The Python loader (brax_rodent_run_b200/mjcf.py + model_blob.py) packs a compiled MJCF model into
 *   dir[2*RR_NFIELDS]  (offset, count) per field, offsets into idata (RR_I) or fdata (RR_F)
 *   idata[]            int32
 *   fdata[]            float64
 * and both consumers -- the CPU oracle (oracle/rr_oracle.c) and the CUDA library (csrc/) -- index it
 * through the enum generated from this list.  model_blob.py parses this header, so adding a field here
 * is the only step needed to make it travel.
 *
 * It replaces the `brax.System` / `mjx.Model` pytree that the reference builds with
 * mjcf_brax.load_model (Rodent_Env_Brax.py:51).  Names follow mjModel.
 */
#ifndef RR_MODEL_FIELDS_H_
#define RR_MODEL_FIELDS_H_

/* integer option vector (field opt_i) */
enum {
  RR_OI_NQ = 0, RR_OI_NV, RR_OI_NU, RR_OI_NA, RR_OI_NBODY, RR_OI_NJNT, RR_OI_NGEOM, RR_OI_NM,
  RR_OI_NPAIR, RR_OI_NCON, RR_OI_NLIMIT, RR_OI_NEFC,
  RR_OI_SOLVER,        /* 0 = CG, 1 = Newton (Rodent_Env_Brax.py:42-45) */
  RR_OI_ITERATIONS,    /* mj_model.opt.iterations     (Rodent_Env_Brax.py:46) */
  RR_OI_LS_ITERATIONS, /* mj_model.opt.ls_iterations  (Rodent_Env_Brax.py:47) */
  RR_OI_COUNT
};
/* float option vector (field opt_f) */
enum {
  RR_OF_TIMESTEP = 0, RR_OF_GRAVITY_X, RR_OF_GRAVITY_Y, RR_OF_GRAVITY_Z, RR_OF_TOLERANCE, RR_OF_LS_TOLERANCE,
  RR_OF_IMPRATIO, RR_OF_MEANINERTIA,
  RR_OF_COUNT
};

#define RR_MODEL_FIELDS(I, F)                                                                         \
  I(opt_i) F(opt_f)                                                                                   \
  /* bodies */                                                                                        \
  I(body_parentid) I(body_rootid) I(body_jntadr) I(body_jntnum) I(body_dofadr) I(body_dofnum)         \
  I(body_lastdof) I(body_subtreesize)                                                                 \
  F(body_pos) F(body_quat) F(body_ipos) F(body_iquat) F(body_inertia) F(body_mass)                    \
  F(body_subtreemass) F(body_invweight0)                                                              \
  /* joints */                                                                                        \
  I(jnt_type) I(jnt_bodyid) I(jnt_qposadr) I(jnt_dofadr) I(jnt_limited)                               \
  F(jnt_pos) F(jnt_axis) F(jnt_range) F(jnt_stiffness) F(jnt_margin) F(jnt_solref) F(jnt_solimp)      \
  F(qpos0) F(qpos_spring)                                                                             \
  /* dofs */                                                                                          \
  I(dof_bodyid) I(dof_jntid) I(dof_parentid)                                                          \
  F(dof_armature) F(dof_damping) F(dof_invweight0)                                                    \
  /* tree-sparse mass matrix layout: row i = ancestors of dof i ascending, diagonal last */           \
  I(M_rowadr) I(M_rownnz) I(M_colind)                                                                 \
  /* geoms */                                                                                         \
  I(geom_type) I(geom_bodyid) F(geom_pos) F(geom_quat) F(geom_size)                                   \
  /* actuators (joint transmission, general gain/bias, none|filter dynamics) */                       \
  I(actuator_jntid) I(actuator_dyntype) I(actuator_gaintype) I(actuator_biastype)                     \
  I(actuator_ctrllimited) I(actuator_forcelimited) I(actuator_actadr)                                 \
  F(actuator_gear) F(actuator_dynprm) F(actuator_gainprm) F(actuator_biasprm)                         \
  F(actuator_ctrlrange) F(actuator_forcerange)                                                        \
  /* static collision-pair table (MJX collision_driver order) + mixed contact parameters */           \
  I(pair_fn) I(pair_geom1) I(pair_geom2) I(pair_conadr)                                               \
  F(pair_friction) F(pair_solref) F(pair_solimp) F(pair_includemargin)                                \
  /* joint-limit constraint rows */                                                                   \
  I(limit_jntid)

#define RR__ENUM_I(n) RR_FIELD_##n,
#define RR__ENUM_F(n) RR_FIELD_##n,
enum { RR_MODEL_FIELDS(RR__ENUM_I, RR__ENUM_F) RR_NFIELDS };

/* collision function ids in pair_fn */
enum { RR_PAIR_PLANE_SPHERE = 0, RR_PAIR_PLANE_CAPSULE = 1, RR_PAIR_PLANE_ELLIPSOID = 2 };
/* mjtJoint */
enum { RR_JNT_FREE = 0, RR_JNT_BALL = 1, RR_JNT_SLIDE = 2, RR_JNT_HINGE = 3 };

#define RR_BLOB_MAGIC 0x52524D31 /* "RRM1" */

#endif /* RR_MODEL_FIELDS_H_ */
