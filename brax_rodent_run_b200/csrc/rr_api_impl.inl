/* rr_api_impl.inl -- body of the C ABI declared in include/rr_b200.h.
 *
 * Included once by rr_api.cu (the product: CUDA backend) and once by tests/emu/rr_emu.cpp (test-only host
 * backend that runs the same kernel text with the 32 lanes as fibers).  The including file provides:
 *   int  rrb_set_device(int device);
 *   int  rrb_malloc(void **p, size_t bytes);   void rrb_free(void *p);
 *   int  rrb_h2d(void *dst, const void *src, size_t bytes, void *stream);
 *   int  rrb_d2h(void *dst, const void *src, size_t bytes, void *stream);
 *   int  rrb_sync(void *stream);
 *   void *rrb_host_devptr(void *host);        (device alias of a pinned host pointer, or null)
 *   int  rrb_geometry(const RRModelDev &m, int B, int *ctas, int *wpb);
 *   int  rrb_num_slots();                      (upper bound on concurrently resident warps = scratch slots)
 *   int  rrb_launch_step(const RRModelDev &m, const RRStepArgs &a, void *stream);
 *   int  rrb_launch_gae(...);
 *   const char *rrb_error();
 */
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/rr_b200.h"
#include "rr_model_build.h"
#include "rr_ppo_loss.h"
#include "rr_learner_misc.h"

static thread_local std::string g_rr_error;
static std::atomic<long long> g_rr_launches{0};

struct rr_model {
  RRHostModel host;
  RRModelDev dev;  /* scalars, layout and table offsets; the table POINTERS are per environment (one copy per device) */
};

struct rr_env {
  const rr_model *model;
  int B, device;
  int32_t *d_ibuf; /* this environment's device copy of the model tables (an rr_model may serve several devices) */
  float *d_fbuf;
  RRTask task;
  float *d_track;
  int episode_length;
  float *d_dbg;
  long long *d_prof;
  float *d_action_stage; /* [B, nu] staging for the host-buffer entry point */
  float *d_scratch;      /* [warp slots, scratch_stride] overflow for contact Jacobians / constraint rows */
  int scratch_stride;
};

static int rr_fail(int code, const std::string &msg) {
  g_rr_error = msg;
  return code;
}

extern "C" const char *rr_last_error(void) { return g_rr_error.c_str(); }
extern "C" long long rr_launch_count(void) { return g_rr_launches.load(); }

extern "C" int rr_model_create(const int32_t *dir, int32_t ndir, const int32_t *idata, int32_t ni, const double *fdata,
                               int32_t nf, rr_model **out) {
  if (!dir || !idata || !fdata || !out) return rr_fail(RR_EINVAL, "rr_model_create: null argument");
  if (ndir != 2 * RR_NFIELDS) return rr_fail(RR_EINVAL, "rr_model_create: blob directory does not match rr_model_fields.h");
  for (int k = 0; k < RR_NFIELDS; k++)
    if (dir[2 * k] < 0 || dir[2 * k + 1] < 0) return rr_fail(RR_EINVAL, "rr_model_create: negative offset in blob directory");
  rr_model *m = new rr_model();
  try {
    rr_host_model_build(m->host, dir, idata, ni, fdata, nf);
  } catch (const std::exception &ex) {
    std::string msg = ex.what();
    delete m;
    return rr_fail(msg.rfind("NotImplemented", 0) == 0 ? RR_ENOTIMPL : RR_EINVAL, msg);
  }
  m->dev = m->host.dev;
  rr_host_model_bind(m->host, m->dev, nullptr, nullptr); /* table offsets; pointers are bound per environment */
  *out = m;
  return RR_OK;
}

extern "C" void rr_model_destroy(rr_model *m) {
  if (!m) return;
  delete m;
}

extern "C" int rr_model_set_solver(rr_model *m, int32_t solver, int32_t iterations, int32_t ls_iterations) {
  if (!m) return rr_fail(RR_EINVAL, "rr_model_set_solver: null model");
  if (solver != 0) return rr_fail(RR_ENOTIMPL, "NotImplemented: only the CG solver is available on the device path");
  if (iterations < 1 || ls_iterations < 0) return rr_fail(RR_EINVAL, "rr_model_set_solver: bad iteration counts");
  m->dev.solver = m->host.dev.solver = solver;
  m->dev.iterations = m->host.dev.iterations = iterations;
  m->dev.ls_iterations = m->host.dev.ls_iterations = ls_iterations;
  return RR_OK;
}

static int rr_debug_stride_of(const RRModelDev &d) {
  int n = 0;
  for (int f = 0; f < rr::RR_DBG_NFIELDS; f++) n += rr::dbg_count(d, f);
  return n;
}

extern "C" int rr_model_dims(const rr_model *m, rr_dims *o) {
  if (!m || !o) return rr_fail(RR_EINVAL, "rr_model_dims: null argument");
  const RRModelDev &d = m->dev;
  o->nq = d.nq; o->nv = d.nv; o->nu = d.nu; o->na = d.na; o->nbody = d.nbody; o->njnt = d.njnt; o->ngeom = d.ngeom;
  o->ncon = d.ncon; o->nlimit = d.nlimit; o->nefc = d.nefc; o->nM = d.nM; o->nroot = d.nroot;
  o->obs_dim = m->host.obs_dim;
  o->smem_bytes = d.sm.total * (int)sizeof(float);
  o->debug_stride = rr_debug_stride_of(d);
  o->timestep = d.timestep;
  return RR_OK;
}

static const char *const kDbgNames[rr::RR_DBG_NFIELDS] = {
    "xpos", "xquat", "subtree_com", "cinert", "cdof", "cvel", "qM_sparse", "qLD_sparse", "qfrc_bias", "qfrc_passive",
    "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "contact_dist", "contact_pos", "contact_frame", "efc_J", "efc_D",
    "efc_aref", "efc_force", "qacc", "qfrc_constraint", "scalars"};

extern "C" int rr_debug_field(const rr_model *m, const char *name, int32_t *offset, int32_t *count) {
  if (!m || !name) return rr_fail(RR_EINVAL, "rr_debug_field: null argument");
  for (int f = 0; f < rr::RR_DBG_NFIELDS; f++)
    if (!std::strcmp(name, kDbgNames[f])) {
      if (offset) *offset = rr::dbg_offset(m->dev, f);
      if (count) *count = rr::dbg_count(m->dev, f);
      return RR_OK;
    }
  return rr_fail(RR_EINVAL, std::string("rr_debug_field: unknown field ") + name);
}

static const char *const kProfNames[RR_NPROF] = {"load", "kinematics", "com_pos", "solver_cost_JTf", "mass_matrix", "factor", "solver_Minv_grad",
                                                  "rne", "smooth", "collision", "make_constraint", "solver_init",
                                                  "solver_linesearch", "solver_update", "euler", "epilogue",
                                                  "wait_substep", "wait_factor", "mul_m_warm", "factor2", "ls_pre", "ls_mul_j", "ls_eval_first2", "ls_loop", "solver_cost"};
extern "C" int rr_prof_count(void) { return RR_NPROF; }
extern "C" const char *rr_prof_name(int32_t i) { return (i >= 0 && i < RR_NPROF) ? kProfNames[i] : ""; }

extern "C" int rr_env_create(const rr_model *cm, int32_t num_envs, int32_t device, rr_env **out) {
  if (!cm || !out || num_envs < 1) return rr_fail(RR_EINVAL, "rr_env_create: bad argument");
  const rr_model *m = cm;
  if (rrb_set_device(device)) return rr_fail(RR_ECUDA, rrb_error());
  rr_env *e = new rr_env();
  std::memset(e, 0, sizeof(*e));
  /* the model tables are uploaded per environment, on the environment's device */
  if (rrb_malloc((void **)&e->d_ibuf, m->host.ibuf.size() * sizeof(int32_t)) ||
      rrb_malloc((void **)&e->d_fbuf, m->host.fbuf.size() * sizeof(float)) ||
      rrb_h2d(e->d_ibuf, m->host.ibuf.data(), m->host.ibuf.size() * sizeof(int32_t), nullptr) ||
      rrb_h2d(e->d_fbuf, m->host.fbuf.data(), m->host.fbuf.size() * sizeof(float), nullptr) || rrb_sync(nullptr)) {
    if (e->d_ibuf) rrb_free(e->d_ibuf);
    if (e->d_fbuf) rrb_free(e->d_fbuf);
    delete e;
    return rr_fail(RR_ECUDA, rrb_error());
  }
  e->model = m;
  e->B = num_envs;
  e->device = device;
  e->task.ctrl_cost_weight = 0.1f;
  e->task.healthy_reward = 1.0f;
  e->task.healthy_z_lo = 0.03f;
  e->task.healthy_z_hi = 0.5f;
  e->task.terminate_when_unhealthy = 1;
  if (rrb_malloc((void **)&e->d_action_stage, (size_t)num_envs * (m->dev.nu > 0 ? m->dev.nu : 1) * sizeof(float))) {
    rrb_free(e->d_ibuf); rrb_free(e->d_fbuf);
    delete e;
    return rr_fail(RR_ECUDA, rrb_error());
  }
  e->scratch_stride = 5 * ((m->dev.nefc + 3) & ~3) + 8 + 6 * ((m->dev.ncon + 3) & ~3) + 32; /* rows (5 nefc) + per-contact six-vectors + profile sums */
  if (rrb_malloc((void **)&e->d_scratch, (size_t)rrb_num_slots() * e->scratch_stride * sizeof(float))) {
    rrb_free(e->d_action_stage); rrb_free(e->d_ibuf); rrb_free(e->d_fbuf);
    delete e;
    return rr_fail(RR_ECUDA, rrb_error());
  }
  *out = e;
  return RR_OK;
}

extern "C" void rr_env_destroy(rr_env *e) {
  if (!e) return;
  if (e->d_scratch) rrb_free(e->d_scratch);
  if (e->d_track) rrb_free(e->d_track);
  if (e->d_action_stage) rrb_free(e->d_action_stage);
  if (e->d_ibuf) rrb_free(e->d_ibuf);
  if (e->d_fbuf) rrb_free(e->d_fbuf);
  delete e;
}

extern "C" int rr_env_set_task(rr_env *e, const float *track_pos, int32_t track_len, float ctrl_cost_weight,
                               float healthy_reward, float healthy_z_lo, float healthy_z_hi, int32_t terminate_when_unhealthy) {
  if (!e || !track_pos || track_len < 1) return rr_fail(RR_EINVAL, "rr_env_set_task: bad argument");
  if (rrb_set_device(e->device)) return rr_fail(RR_ECUDA, rrb_error());
  if (e->d_track) rrb_free(e->d_track);
  e->d_track = nullptr;
  if (rrb_malloc((void **)&e->d_track, (size_t)track_len * 3 * sizeof(float)) ||
      rrb_h2d(e->d_track, track_pos, (size_t)track_len * 3 * sizeof(float), nullptr) || rrb_sync(nullptr))
    return rr_fail(RR_ECUDA, rrb_error());
  e->task.track_pos = e->d_track;
  e->task.track_len = track_len;
  e->task.ctrl_cost_weight = ctrl_cost_weight;
  e->task.healthy_reward = healthy_reward;
  e->task.healthy_z_lo = healthy_z_lo;
  e->task.healthy_z_hi = healthy_z_hi;
  e->task.terminate_when_unhealthy = terminate_when_unhealthy ? 1 : 0;
  return RR_OK;
}

extern "C" int rr_env_set_wrappers(rr_env *e, int32_t episode_length) {
  if (!e || episode_length < 0) return rr_fail(RR_EINVAL, "rr_env_set_wrappers: bad argument");
  e->episode_length = episode_length;
  return RR_OK;
}

extern "C" int rr_env_geometry(const rr_env *e, int32_t *ctas, int32_t *envs_per_cta, int32_t *passes) {
  if (!e) return rr_fail(RR_EINVAL, "rr_env_geometry: null env");
  int g = 1, w = 1;
  if (rrb_set_device(e->device) || rrb_geometry(e->model->dev, e->B, &g, &w)) return rr_fail(RR_ECUDA, rrb_error());
  if (ctas) *ctas = g;
  if (envs_per_cta) *envs_per_cta = w;
  if (passes) *passes = (e->B + g * w - 1) / (g * w);
  return RR_OK;
}

extern "C" int rr_env_set_debug(rr_env *e, float *dbg) {
  if (!e) return rr_fail(RR_EINVAL, "rr_env_set_debug: null env");
  e->d_dbg = dbg;
  return RR_OK;
}
extern "C" int rr_env_set_profile(rr_env *e, long long *prof) {
  if (!e) return rr_fail(RR_EINVAL, "rr_env_set_profile: null env");
  e->d_prof = prof;
  return RR_OK;
}

static int rr_fill_args(rr_env *e, const rr_buffers *b, const float *action, int nsub, int mode, RRStepArgs &a) {
  if (!e || !b) return rr_fail(RR_EINVAL, "null env / buffers");
  if (!b->qpos || !b->qvel || !b->qacc_warmstart || (e->model->dev.na > 0 && !b->act))
    return rr_fail(RR_EINVAL, "rr_buffers: qpos, qvel, act and qacc_warmstart are required");
  const bool needs_task = b->obs || b->reward || b->metrics;
  if (needs_task && !e->task.track_pos) return rr_fail(RR_EINVAL, "rr_env_set_task must be called before obs / reward are requested");
  if (needs_task && !b->cur_frame) return rr_fail(RR_EINVAL, "rr_buffers.cur_frame is required with obs / reward");
  std::memset(&a, 0, sizeof(a));
  a.B = e->B; a.nsub = nsub; a.mode = mode; a.action = action;
  a.qpos = b->qpos; a.qvel = b->qvel; a.act = b->act; a.warm = b->qacc_warmstart; a.time = b->time;
  a.cur_frame = b->cur_frame;
  a.in_qpos = b->in_qpos ? b->in_qpos : b->qpos; a.in_qvel = b->in_qvel ? b->in_qvel : b->qvel;
  a.in_act = b->in_act ? b->in_act : b->act; a.in_warm = b->in_qacc_warmstart ? b->in_qacc_warmstart : b->qacc_warmstart;
  a.in_time = b->in_time ? b->in_time : b->time; a.in_cur_frame = b->in_cur_frame ? b->in_cur_frame : b->cur_frame;
  a.in_done = b->in_done ? b->in_done : b->done; a.in_steps = b->in_steps ? b->in_steps : b->steps;
  a.task = e->task;
  a.obs = b->obs; a.reward = b->reward; a.done = b->done; a.metrics = b->metrics;
  a.wrap = e->episode_length > 0;
  a.episode_length = e->episode_length;
  if (a.wrap) {
    if (!b->done || !b->steps || !b->truncation) return rr_fail(RR_EINVAL, "wrappers need done, steps and truncation buffers");
    if (mode == RR_MODE_STEP && (!b->first_qpos || !b->first_qvel || !b->first_qacc_warmstart ||
                                 (e->model->dev.na > 0 && !b->first_act) || (b->obs && !b->first_obs)))
      return rr_fail(RR_EINVAL, "wrappers need the first_* buffers");
  }
  a.steps = b->steps; a.truncation = b->truncation;
  a.first_qpos = b->first_qpos; a.first_qvel = b->first_qvel; a.first_act = b->first_act;
  a.first_warm = b->first_qacc_warmstart; a.first_time = b->first_time; a.first_obs = b->first_obs;
  a.xpos = b->xpos; a.xquat = b->xquat; a.subtree_com = b->subtree_com; a.qfrc_actuator = b->qfrc_actuator;
  a.cinert = b->cinert; a.cvel = b->cvel; a.contact_dist = b->contact_dist; a.qacc = b->qacc; a.niter = b->solver_niter;
  a.work = b->work; a.env_order = b->env_order;
  a.contact_pos = b->contact_pos; a.contact_frame = b->contact_frame;
  a.dbg.buf = e->d_dbg;
  a.dbg.stride = rr_debug_stride_of(e->model->dev);
  a.prof = e->d_prof;
  a.scratch = e->d_scratch;
  a.scratch_stride = e->scratch_stride;
  return RR_OK;
}

/* the model's scalars / layout with THIS environment's device tables */
static RRModelDev rr_env_dev(const rr_env *e) {
  RRModelDev d = e->model->dev;
  d.ibuf = e->d_ibuf;
  d.fbuf = e->d_fbuf;
  return d;
}

extern "C" int rr_env_init(rr_env *e, const rr_buffers *b, void *stream) {
  RRStepArgs a;
  int rc = rr_fill_args(e, b, nullptr, 0, RR_MODE_INIT, a);
  if (rc) return rc;
  if (rrb_set_device(e->device) || rrb_launch_step(rr_env_dev(e), a, stream)) return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches++;
  return RR_OK;
}

extern "C" int rr_env_step(rr_env *e, const rr_buffers *b, const float *action, int32_t n_frames, void *stream) {
  if (n_frames < 1) return rr_fail(RR_EINVAL, "rr_env_step: n_frames must be >= 1");
  if (!action) return rr_fail(RR_EINVAL, "rr_env_step: null action");
  RRStepArgs a;
  int rc = rr_fill_args(e, b, action, n_frames, RR_MODE_STEP, a);
  if (rc) return rc;
  if (rrb_set_device(e->device) || rrb_launch_step(rr_env_dev(e), a, stream)) return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches++;
  return RR_OK;
}

extern "C" int rr_env_step_host(rr_env *e, const rr_buffers *b, const float *action_host, int32_t n_frames, float *obs_host,
                                float *reward_host, float *done_host, void *stream) {
  if (!e || !b || !action_host) return rr_fail(RR_EINVAL, "rr_env_step_host: null argument");
  const RRModelDev &d = e->model->dev;
  if (rrb_set_device(e->device) || rrb_h2d(e->d_action_stage, action_host, (size_t)e->B * d.nu * sizeof(float), stream))
    return rr_fail(RR_ECUDA, rrb_error());
  /* Pinned (device-accessible) obs_host: the kernel stores the observation -- 97 % of the bytes that leave the device --
   * straight into it over PCIe while it computes, instead of a 20 MB copy after the kernel.  b->obs is then left untouched. */
  float *obs_direct = obs_host && b->obs ? (float *)rrb_host_devptr(obs_host) : nullptr;
  rr_buffers bb = *b;
  if (obs_direct) bb.obs = obs_direct;
  int rc = rr_env_step(e, &bb, e->d_action_stage, n_frames, stream);
  if (rc) return rc;
  if (!obs_direct && obs_host && b->obs && rrb_d2h(obs_host, b->obs, (size_t)e->B * e->model->host.obs_dim * sizeof(float), stream))
    return rr_fail(RR_ECUDA, rrb_error());
  if (reward_host && b->reward && rrb_d2h(reward_host, b->reward, (size_t)e->B * sizeof(float), stream))
    return rr_fail(RR_ECUDA, rrb_error());
  if (done_host && b->done && rrb_d2h(done_host, b->done, (size_t)e->B * sizeof(float), stream))
    return rr_fail(RR_ECUDA, rrb_error());
  if (rrb_sync(stream)) return rr_fail(RR_ECUDA, rrb_error());
  return RR_OK;
}

extern "C" int rr_gae(const float *rewards, const float *values, const float *bootstrap_value, const float *termination,
                      const float *truncation, int32_t T, int32_t B, float discount, float lambda_, float *vs,
                      float *advantages, void *stream) {
  if (!rewards || !values || !bootstrap_value || !termination || !truncation || !vs || !advantages || T < 1 || B < 1)
    return rr_fail(RR_EINVAL, "rr_gae: bad argument");
  if (rrb_launch_gae(rewards, values, bootstrap_value, termination, truncation, T, B, discount, lambda_, vs, advantages, stream))
    return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches++;
  return RR_OK;
}


extern "C" int rr_ppo_loss_blocks(int32_t T, int32_t B, int32_t *blocks_a, int32_t *blocks_b) {
  if (T < 1 || B < 1 || !blocks_a || !blocks_b) return rr_fail(RR_EINVAL, "rr_ppo_loss_blocks: bad argument");
  *blocks_a = rrb_ppo_blocks(B);
  *blocks_b = rrb_ppo_blocks_b(T * B);
  return RR_OK;
}

extern "C" int rr_ppo_loss(const rr_ppo_loss_args *u, void *stream) {
  if (!u) return rr_fail(RR_EINVAL, "rr_ppo_loss: null argument");
  if (u->T < 1 || u->B < 1 || u->A < 1 || !u->logits || !u->baseline || !u->bootstrap || !u->raw_action || !u->old_log_prob ||
      !u->reward || !u->discount || !u->truncation || !u->noise || !u->scratch || !u->adv_partial || !u->loss_partial ||
      !u->grad_logits || !u->grad_baseline)
    return rr_fail(RR_EINVAL, "rr_ppo_loss: bad argument");
  RRPpoLossArgs a;
  a.T = u->T; a.B = u->B; a.A = u->A;
  a.logits = u->logits; a.baseline = u->baseline; a.bootstrap = u->bootstrap; a.raw_action = u->raw_action;
  a.old_log_prob = u->old_log_prob; a.reward = u->reward; a.discount = u->discount; a.truncation = u->truncation; a.noise = u->noise;
  a.reward_scaling = u->reward_scaling; a.gamma = u->discounting; a.lambda_ = u->gae_lambda; a.clip_eps = u->clipping_epsilon;
  a.entropy_cost = u->entropy_cost; a.normalize_advantage = u->normalize_advantage;
  const size_t n = (size_t)u->T * u->B;
  a.lp = u->scratch; a.adv = u->scratch + n; a.vs = u->scratch + 2 * n;
  a.adv_partial = u->adv_partial; a.loss_partial = u->loss_partial;
  a.grad_logits = u->grad_logits; a.grad_baseline = u->grad_baseline;
  a.nblkA = rrb_ppo_blocks(u->B);
  if (rrb_launch_ppo_loss(a, stream)) return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches += 2;
  return RR_OK;
}

/* ---- grouped tensor-core GEMM of the learner ------------------------------------------------------------------ */
extern "C" int32_t rr_tc_record_bytes(void) { return (int32_t)sizeof(RRTcRecord); }

extern "C" int rr_tc_plan(rr_tc_problem *pr, int32_t count, int32_t *total_tiles, int32_t *smem_bytes, void *records) {
  if (!pr || count < 1 || count > 256 || !total_tiles || !smem_bytes || !records) return rr_fail(RR_EINVAL, "rr_tc_plan: bad argument");
  int tiles = 0;
  for (int i = 0; i < count; i++) {
    rr_tc_problem &p = pr[i];
    if (p.m < 1 || p.n < 1 || p.k < 1 || !p.a || !p.b || !p.d) return rr_fail(RR_EINVAL, "rr_tc_plan: empty problem or null matrix");
    if (p.epi < 0 || p.epi > 2 || (p.epi == 2 && !p.aux_in)) return rr_fail(RR_EINVAL, "rr_tc_plan: bad epilogue");
    if (p.b_ones < 0 || p.b_ones > 2 || (p.b_ones && !p.ones_out) || (p.b_ones == 1 && !p.b_mn) || (p.b_ones == 2 && p.n < 2))
      return rr_fail(RR_EINVAL, "rr_tc_plan: b_ones needs ones_out (and an MN-major B for the virtual row)");
    const int n_d = p.n - (p.b_ones == 2 ? 1 : 0);
    if (p.lda < (p.a_mn ? p.m : p.k) || p.ldb < (p.b_mn ? p.n : p.k) || p.ldd < n_d ||
        ((p.epi == 2 || (p.epi == 1 && p.aux_out)) && p.ldaux < n_d))
      return rr_fail(RR_EINVAL, "rr_tc_plan: leading dimension smaller than the row length");
    const int n_ext = n_d + (p.b_ones ? 1 : 0), tiles_m = (p.m + 127) / 128;
    /* 128-wide tiles halve the A traffic and the tensor core's shared-memory reads per flop; narrower only when n is */
    int bn = (n_ext + 15) / 16 * 16;
    if (bn > 128) bn = 128;
    p.bn = bn;
    p.tiles_n = (n_ext + bn - 1) / bn;
    p.tile_start = tiles;
    tiles += tiles_m * p.tiles_n;
  }
  *total_tiles = tiles;
  /* shared memory of the launch: the largest ring; a problem's ring is 4 stages (3 for the 128-wide tile: two CTAs per SM) of
   * one 32-wide k-block of A (16 KB) and B */
  int smem = 0;
  for (int i = 0; i < count; i++) {
    rr_tc_problem &p = pr[i];
    const int stage = 128 * 128 + (p.b_mn ? (p.bn + 31) / 32 * 4096 : p.bn * 128);
    const int nst = p.bn > 64 ? 3 : 4;
    p.reserved[3] = nst;
    if (nst * stage > smem) smem = nst * stage;
  }
  *smem_bytes = smem;
  if (*smem_bytes > rrb_tc_smem_max()) return rr_fail(RR_EINVAL, "rr_tc_plan: tile does not fit shared memory");
  /* device records: the planned problem + the tensor maps of the operands TMA can fetch (backend-specific; none on the emulator) */
  for (int i = 0; i < count; i++) {
    RRTcRecord rec;
    memset(&rec, 0, sizeof(rec));
    pr[i].reserved[2] = 0;
    rec.p = pr[i];
    rrb_tc_encode(rec);
    pr[i].reserved[2] = rec.p.reserved[2];
    memcpy((char *)records + (size_t)i * sizeof(RRTcRecord), &rec, sizeof(rec));
  }
  return RR_OK;
}

extern "C" int rr_tc_launch(const void *dev_problems, int32_t count, int32_t total_tiles, int32_t smem_bytes, void *stream) {
  if (!dev_problems || count < 1 || total_tiles < 1 || smem_bytes < 1) return rr_fail(RR_EINVAL, "rr_tc_launch: bad argument");
  if (rrb_tc_launch((const RRTcRecord *)dev_problems, count, total_tiles, smem_bytes, stream)) return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches += 1;
  return RR_OK;
}

extern "C" int rr_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *step, int64_t n, float lr,
                            float beta1, float beta2, float eps, void *stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step || n < 1) return rr_fail(RR_EINVAL, "rr_adam_step: bad argument");
  if (rrb_adam_step(param, const_cast<float *>(grad), nullptr, 0, exp_avg, exp_avg_sq, step, n, lr, beta1, beta2, eps, stream))
    return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches += 2;
  return RR_OK;
}

extern "C" int rr_adam_step_sum(float *param, float *grad, const float *partials, int32_t nsplit, float *exp_avg, float *exp_avg_sq,
                                float *step, int64_t n, float lr, float beta1, float beta2, float eps, void *stream) {
  if (!param || !grad || !partials || nsplit < 1 || !exp_avg || !exp_avg_sq || !step || n < 1)
    return rr_fail(RR_EINVAL, "rr_adam_step_sum: bad argument");
  if (rrb_adam_step(param, grad, partials, nsplit, exp_avg, exp_avg_sq, step, n, lr, beta1, beta2, eps, stream))
    return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches += 2;
  return RR_OK;
}

extern "C" int rr_gather_rows(const rr_gather_item *items, int32_t count, const int64_t *idx, int32_t rows, void *stream) {
  if (!items || count < 1 || count > RR_GATHER_MAX || !idx || rows < 1) return rr_fail(RR_EINVAL, "rr_gather_rows: bad argument");
  RRGatherArgs a;
  memset(&a, 0, sizeof(a));
  a.count = count; a.rows = rows; a.idx = idx;
  int blocks = 0;
  for (int i = 0; i < count; i++) {
    if (!items[i].src || !items[i].dst || items[i].outer < 1 || items[i].src_rows < 1 || items[i].inner < 1)
      return rr_fail(RR_EINVAL, "rr_gather_rows: bad item");
    a.item[i] = items[i];
    if (a.item[i].dst_pitch == 0) a.item[i].dst_pitch = items[i].inner;
    if (a.item[i].dst_pitch < items[i].inner) return rr_fail(RR_EINVAL, "rr_gather_rows: dst_pitch smaller than the row");
    a.block_start[i] = blocks;
    blocks += items[i].outer * rows;
  }
  a.block_start[count] = blocks;
  if (rrb_gather_rows(a, blocks, stream)) return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches += 1;
  return RR_OK;
}

extern "C" int rr_policy_act(const rr_policy_args *a, void *stream) {
  if (!a || !a->obs || !a->action || !a->raw_action || !a->log_prob || a->B < 1 || a->obs_dim < 1 || a->in0 < a->obs_dim ||
      a->nlayers < 2 || a->nlayers > RR_POLICY_MAX_LAYERS || a->A < 1 || a->A > 32 || (a->mean && !a->std))
    return rr_fail(RR_EINVAL, "rr_policy_act: bad argument (hidden layers are 32 wide, at most 32 actions)");
  for (int l = 0; l < a->nlayers; l++)
    if (!a->w[l] || !a->b[l]) return rr_fail(RR_EINVAL, "rr_policy_act: null layer");
  if (rrb_policy_act(*a, stream)) return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches += 1;
  return RR_OK;
}

extern "C" int rr_measure_fp32_peak(double *tflops, void *stream) {
  if (!tflops) return rr_fail(RR_EINVAL, "rr_measure_fp32_peak: null argument");
  if (rrb_fp32_peak(tflops, stream)) return rr_fail(RR_ECUDA, rrb_error());
  g_rr_launches += 4;
  return RR_OK;
}
