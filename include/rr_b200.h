/* rr_b200.h -- C ABI of the B200 rodent step library (librr_b200.so).
 *
 * This is the drop-in boundary for the reference's data-parallel hot path.  Each entry point names the
 * reference interface it replaces (file:line in talmolab/Brax-Rodent-Run); the Python class
 * brax_rodent_run_b200.Rodent binds these with ctypes and mirrors Rodent_Env_Brax.py's surface.
 * Plain pointers and sizes only; all `float *` / `int32_t *` members of rr_buffers are DEVICE pointers
 * into caller-owned, env-major (row per environment) arrays; work is enqueued on the caller's CUDA
 * stream (`stream` is a cudaStream_t passed as void *) and nothing synchronises unless stated.
 *
 * Return value: 0 on success, otherwise an RR_E* code; rr_last_error() gives the message.
 *   RR_EINVAL  -> Python raises ValueError        (bad arguments / bad model blob)
 *   RR_ENOTIMPL-> Python raises NotImplementedError (model feature outside the supported subset, as
 *                 brax.io.mjcf.load_model does, Rodent_Env_Brax.py:51)
 *   RR_ECUDA   -> Python raises RuntimeError
 */
#ifndef RR_B200_H_
#define RR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { RR_OK = 0, RR_EINVAL = 1, RR_ENOTIMPL = 2, RR_ECUDA = 3 };

typedef struct rr_model rr_model;
typedef struct rr_env rr_env;

typedef struct rr_dims {
  int32_t nq, nv, nu, na, nbody, njnt, ngeom, ncon, nlimit, nefc, nM, nroot;
  int32_t obs_dim;       /* nq + nv + 10 (nbody-1) + 6 (nbody-1) + nv + 3, Rodent_Env_Brax.py:149-158 */
  int32_t smem_bytes;    /* shared memory per environment (one warp) */
  int32_t debug_stride;  /* floats per environment of the debug record */
  float timestep;        /* mj_model.opt.timestep */
} rr_dims;

/* Caller-owned device arrays, one row per environment.  Null members are skipped where marked optional. */
typedef struct rr_buffers {
  /* persistent physics state = the fields of mjx.Data that the next step reads */
  float *qpos;            /* [B, nq]  */
  float *qvel;            /* [B, nv]  */
  float *act;             /* [B, na]  */
  float *qacc_warmstart;  /* [B, nv]  */
  float *time;            /* [B] optional */
  int32_t *cur_frame;     /* [B] state.info["cur_frame"], Rodent_Env_Brax.py:77-79,103-104 */
  /* Optional separate INPUT arrays (functional style, as the reference's immutable State): when non-null the
   * step reads the previous state from these and writes the new one to the members above; when null the
   * members above are updated in place. */
  const float *in_qpos, *in_qvel, *in_act, *in_qacc_warmstart, *in_time;
  const int32_t *in_cur_frame;
  const float *in_done, *in_steps; /* previous done / steps for the fused wrappers */
  /* step outputs (State fields) */
  float *obs;             /* [B, obs_dim] optional */
  float *reward;          /* [B] optional */
  float *done;            /* [B] optional; required (read as previous done) when wrappers are on */
  float *metrics;         /* [B, 3] pos_reward, reward_quadctrl, reward_alive; optional */
  /* brax EpisodeWrapper / AutoResetWrapper state (required when rr_env_set_wrappers(len > 0)) */
  float *steps;           /* [B] info["steps"] */
  float *truncation;      /* [B] info["truncation"] */
  const float *first_qpos, *first_qvel, *first_act, *first_qacc_warmstart, *first_time, *first_obs;
  /* optional views of the last forward pass (pipeline_state attributes) */
  float *xpos;            /* [B, nbody, 3] */
  float *xquat;           /* [B, nbody, 4] */
  float *subtree_com;     /* [B, nroot, 3] */
  float *qfrc_actuator;   /* [B, nv] */
  float *cinert;          /* [B, nbody, 10] */
  float *cvel;            /* [B, nbody, 6] */
  float *contact_dist;    /* [B, ncon] */
  float *qacc;            /* [B, nv] */
  int32_t *solver_niter;  /* [B] */
  /* Load balancing (optional).  work[B]: out, CG iterations this environment's solver ran over the substeps of the step (a cost
   * estimate for the next step).  env_order: in, slot -> environment index (-1 = idle slot), length rr_env_num_slots(); the warps of
   * a CTA rendezvous every substep, so a caller that groups environments of similar cost into the same CTA (see
   * Rodent._balance) shortens the wait.  Null = identity order. */
  float *work;
  const int32_t *env_order;
  /* contact views of the last forward pass (brax.mjx State.contact: torchrl_explore.ipynb:609-618), optional */
  float *contact_pos;     /* [B, ncon, 3] world */
  float *contact_frame;   /* [B, ncon, 3, 3] rows = normal, tangent 1, tangent 2 */
} rr_buffers;

const char *rr_last_error(void);

/* Model: replaces mujoco.MjModel.from_xml_path + brax.io.mjcf.load_model (Rodent_Env_Brax.py:41,51).
 * Input is the flat-model blob of include/rr_model_fields.h produced by brax_rodent_run_b200.mjcf. */
int rr_model_create(const int32_t *dir, int32_t ndir, const int32_t *idata, int32_t ni, const double *fdata,
                    int32_t nf, rr_model **out);
void rr_model_destroy(rr_model *m);
/* mj_model.opt.solver / iterations / ls_iterations (Rodent_Env_Brax.py:42-47); solver: 0 = CG */
int rr_model_set_solver(rr_model *m, int32_t solver, int32_t iterations, int32_t ls_iterations);
int rr_model_dims(const rr_model *m, rr_dims *out);

/* Environment batch on one GPU: replaces PipelineEnv.__init__ (Rodent_Env_Brax.py:60) + VmapWrapper. */
int rr_env_create(const rr_model *m, int32_t num_envs, int32_t device, rr_env **out);
void rr_env_destroy(rr_env *e);
/* Run-task constants (Rodent_Env_Brax.py:62-69); track_pos is a HOST array [track_len, 3], copied. */
int rr_env_set_task(rr_env *e, const float *track_pos, int32_t track_len, float ctrl_cost_weight, float healthy_reward,
                    float healthy_z_lo, float healthy_z_hi, int32_t terminate_when_unhealthy);
/* brax EpisodeWrapper(episode_length, action_repeat=1) + AutoResetWrapper fused into the step; 0 = off */
int rr_env_set_wrappers(rr_env *e, int32_t episode_length);
/* Launch geometry of the step kernel for this batch: CTAs, environments (warps) per CTA, passes per warp.  Slot p of
 * env_order is pass * (ctas * envs_per_cta) + cta * envs_per_cta + warp. */
int rr_env_geometry(const rr_env *e, int32_t *ctas, int32_t *envs_per_cta, int32_t *passes);

/* pipeline_init + _get_obs of Rodent.reset (Rodent_Env_Brax.py:87-95): mjx.forward at (qpos, qvel) with
 * act = ctrl = 0 is the caller's job to zero; writes normalised qpos, qacc_warmstart, obs, zero reward/done/metrics. */
int rr_env_init(rr_env *e, const rr_buffers *b, void *stream);
/* Rodent.step (Rodent_Env_Brax.py:98-136): n_frames x mjx.step with ctrl = action, reward, done, metrics, obs
 * (+ the fused wrappers when enabled).  action: DEVICE [B, nu]. */
int rr_env_step(rr_env *e, const rr_buffers *b, const float *action, int32_t n_frames, void *stream);
/* Same step through HOST buffers: copies action host->device, runs the step, brings obs / reward / done to the host and
 * synchronises the stream.  When obs_host is pinned (device-accessible) memory the kernel stores the observation straight
 * into it (zero-copy) and b->obs is not written; pageable memory gets a copy after the kernel.  This is the call the
 * end-to-end benchmark times. */
int rr_env_step_host(rr_env *e, const rr_buffers *b, const float *action_host, int32_t n_frames, float *obs_host,
                     float *reward_host, float *done_host, void *stream);

/* ppo.losses.compute_gae (brax, called inside ppo.train from brax_rodent_run_ppo.py:200).  All DEVICE,
 * time-major: rewards/values/termination/truncation [T, B], bootstrap [B]; outputs vs, advantages [T, B]. */
int rr_gae(const float *rewards, const float *values, const float *bootstrap_value, const float *termination,
           const float *truncation, int32_t T, int32_t B, float discount, float lambda_, float *vs, float *advantages,
           void *stream);

/* Fused PPO minibatch loss + gradient w.r.t. the network outputs.  Replaces, for one minibatch, what brax's
 * ppo.losses.compute_ppo_loss computes under jax.grad inside ppo.train (brax_rodent_run_ppo.py:97-114, 200): tanh-normal
 * log-prob, compute_gae, advantage normalisation, clipped surrogate, 0.25 * mse value loss, sampled entropy.
 * All pointers are device pointers, time-major [T, B, ...] contiguous fp32 (adv_partial: double).
 *   scratch       [3 T B]           work space (log-prob, advantages, vs)
 *   adv_partial   [2 blocks_a]      doubles; loss_partial [3 blocks_b]: per-block sums of (policy, value, entropy) terms --
 *                                   the caller divides their column sums by T B (rr_ppo_loss_blocks gives the block counts)
 *   grad_logits   [T, B, 2 A], grad_baseline [T, B]: d total_loss / d output, total = policy + value - entropy_cost * entropy */
typedef struct rr_ppo_loss_args {
  int32_t T, B, A;
  const float *logits, *baseline, *bootstrap, *raw_action, *old_log_prob, *reward, *discount, *truncation, *noise;
  float reward_scaling, discounting, gae_lambda, clipping_epsilon, entropy_cost;
  int32_t normalize_advantage;
  float *scratch;
  double *adv_partial;
  float *loss_partial, *grad_logits, *grad_baseline;
} rr_ppo_loss_args;
int rr_ppo_loss_blocks(int32_t T, int32_t B, int32_t *blocks_a, int32_t *blocks_b);
int rr_ppo_loss(const rr_ppo_loss_args *args, void *stream);


/* Grouped TF32 tensor-core GEMM with fused epilogues: the contractions of the PPO learner (brax ppo.train's policy / value MLPs
 * under jax.grad, brax_rodent_run_ppo.py:97-114, 200).  One launch runs a LIST of problems D = epi(A B' + bias), A logical
 * [m x k], B logical [n x k], fp32 in memory, TF32 products with fp32 accumulation in tensor memory (tcgen05.mma kind::tf32,
 * one 128 x bn tile per CTA, operands staged by cp.async straight into the UMMA canonical no-swizzle layouts).
 *   a_mn / b_mn   0: K-major, element (r, kk) at r * ld + kk;   1: MN-major, element (r, kk) at kk * ld + r
 *                 (so X W', dY W and dY' X all run without a transposed copy)
 *   epi           0: D = acc + bias
 *                 1: z = acc + bias; aux_out (if given) = z; D = silu(z)
 *                 2: D = (acc + bias) * silu'(aux_in)         (dgrad through the previous layer's activation)
 *   b_ones        1: B gets a virtual extra row n of ones (MN-major B only): column n of the product, the sum of A over k -- the
 *                 bias gradient when A = dY' -- goes to ones_out[m].  2: B's LAST row (index n - 1) is that row of ones, stored in
 *                 memory (the activation buffers carry a column of ones): D has n - 1 columns, column n - 1 goes to ones_out
 * rr_tc_plan validates a HOST array, fills bn / tile_start / tiles_n and writes one device record per problem (the planned problem
 * + TMA tensor maps of the operands whose base and pitch are 16-byte aligned; the others are fetched with cp.async);
 * rr_tc_launch takes a DEVICE copy of the records (so that the launch is capturable in a CUDA graph).  All matrices are DEVICE
 * pointers. */
typedef struct rr_tc_problem {
  const float *a, *b;
  float *d;
  const float *bias, *aux_in;
  float *aux_out, *ones_out;
  int32_t m, n, k, lda, ldb, ldd, ldaux;
  int32_t a_mn, b_mn, epi, b_ones;
  int32_t bn, tile_start, tiles_n; /* filled by rr_tc_plan */
  int32_t reserved[4];
} rr_tc_problem;
int32_t rr_tc_record_bytes(void);
int rr_tc_plan(rr_tc_problem *host_problems, int32_t count, int32_t *total_tiles, int32_t *smem_bytes,
               void *host_records /* count x rr_tc_record_bytes(): copy to the device for rr_tc_launch */);
int rr_tc_launch(const void *device_records, int32_t count, int32_t total_tiles, int32_t smem_bytes, void *stream);

/* Adam on flat fp32 arrays with torch.optim.Adam's arithmetic (no weight decay, no amsgrad; optax.adam in brax's ppo.train):
 * `step` is a DEVICE float holding the number of steps taken so far; the call uses step + 1 in the bias corrections and then
 * stores step + 1 (capturable in a CUDA graph).  All pointers DEVICE. */
int rr_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *step, int64_t n, float lr, float beta1,
                 float beta2, float eps, void *stream);
/* The same with the gradient given as `nsplit` partial sums (partials [nsplit, n] contiguous: the split weight-gradient launch's
 * workspace): grad[i] = sum over s (fixed order) is formed, stored and stepped over in one pass. */
int rr_adam_step_sum(float *param, float *grad, const float *partials, int32_t nsplit, float *exp_avg, float *exp_avg_sq, float *step,
                     int64_t n, float lr, float beta1, float beta2, float eps, void *stream);

/* Minibatch gather of the learner: for every item, dst[t, j, :] = src[t, idx[j], :] (src [outer, src_rows, inner] contiguous,
 * dst [outer, rows, inner] with row pitch dst_pitch, fp32); one launch for all items (at most 8).  idx: DEVICE int64 [rows]. */
typedef struct rr_gather_item {
  const float *src;
  float *dst;
  int32_t outer, src_rows, inner, dst_pitch; /* dst_pitch: floats between dst rows (0 = inner) */
} rr_gather_item;
int rr_gather_rows(const rr_gather_item *host_items, int32_t count, const int64_t *idx, int32_t rows, void *stream);

/* Policy inference of the rollout (brax acting.generate_unroll -> policy apply, brax_rodent_run_ppo.py:97-114): observation
 * normalisation, the policy MLP (hidden layers of width 32, swish) and the tanh-normal sample + log-prob in ONE kernel, fp32.
 * weights: torch.nn.Linear layout [out, in] row-major (layer 0 with row pitch in0 >= obs_dim: zero-padded input columns are
 * skipped); eps: standard normal noise [B, A] (null: deterministic, action = tanh(loc), log_prob = 0).  All pointers DEVICE. */
#define RR_POLICY_MAX_LAYERS 8
typedef struct rr_policy_args {
  const float *obs, *mean, *std;                 /* [B, obs_dim], [obs_dim], [obs_dim] (mean null: no normalisation) */
  const float *w[RR_POLICY_MAX_LAYERS], *b[RR_POLICY_MAX_LAYERS];
  const float *eps;
  float *action, *raw_action, *log_prob;         /* [B, A], [B, A], [B] */
  int32_t B, obs_dim, in0, nlayers, A, reserved[3]; /* nlayers Linear layers: nlayers - 1 hidden (32 wide) + the 2A-wide head */
} rr_policy_args;
int rr_policy_act(const rr_policy_args *args, void *stream);

/* Parity-test hooks: per-environment dump of forward-pass intermediates (tests only). */
int rr_debug_field(const rr_model *m, const char *name, int32_t *offset, int32_t *count);
int rr_env_set_debug(rr_env *e, float *dbg /* DEVICE [B, debug_stride] or null */);
/* Per-phase clock64 accumulation (profiling builds of bench.py); prof: DEVICE int64 [B, rr_prof_count()] or null */
int rr_env_set_profile(rr_env *e, long long *prof);
int rr_prof_count(void);
const char *rr_prof_name(int32_t i);
/* Measured FP32 FMA throughput of the current device (TFLOP/s, 2 flops per FMA): 8 independent FMA chains per thread on
 * every SM for ~10 ms, timed with CUDA events on `stream`.  The denominator of bench.py's FP32 roofline;
 * no counterpart in the reference. */
int rr_measure_fp32_peak(double *tflops, void *stream);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long rr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RR_B200_H_ */
