"""PPO on the B200 rodent env -- the torch counterpart of the `ppo.train` call in brax_rodent_run_ppo.py:97-114,200.

Follows brax 0.10.x `brax.training.agents.ppo` (restated from its published algorithm, brax is
not installable here): policy MLP (32,)*4 -> 2*nu, value MLP (256,)*5 -> 1, swish, lecun-uniform init; tanh-normal
policy with scale = softplus(.) + 1e-3; running mean/std observation normaliser updated once per training step; GAE
(the rr_gae CUDA kernel); clipped surrogate (eps 0.3) + 0.25 * value error^2 - entropy_cost * entropy; Adam; per
training step `batch_size * num_minibatches / num_envs` unrolls of `unroll_length` env steps, then
`num_updates_per_batch` epochs over `num_minibatches` shuffled minibatches.

Multi-GPU: one process per GPU (torchrun), environments sharded per rank with no communication during rollout; NCCL
all-reduce only for the flat gradient (one bucket per minibatch), the normaliser moments (once per training step)
and the metrics -- the `lax.pmean / psum` calls of brax's pmap'd train step.
"""
from __future__ import annotations

import collections.abc
import ctypes
import dataclasses
import math
import time
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .env import Rodent, State


@dataclasses.dataclass
class PPOConfig:
    """Defaults = the arguments of train_fn in brax_rodent_run_ppo.py:97-114 plus brax's own defaults."""
    num_timesteps: int = 10_000_000
    episode_length: int = 1000
    num_envs: int = 2048            # per process
    batch_size: int = 512           # per process
    num_minibatches: int = 64
    unroll_length: int = 10
    num_updates_per_batch: int = 8
    learning_rate: float = 3e-4
    discounting: float = 0.97
    gae_lambda: float = 0.95
    clipping_epsilon: float = 0.3
    entropy_cost: float = 1e-3
    reward_scaling: float = 1.0
    normalize_observations: bool = True
    normalize_advantage: bool = True
    num_evals: int = 1
    num_eval_envs: int = 128
    deterministic_eval: bool = False
    seed: int = 0
    tc_learner: Optional[bool] = None  # the two MLPs' forward / backward as 12 grouped TMA + tcgen05 GEMM launches (tc_learner.py),
                                    # Adam and the minibatch gather as own kernels, instead of autograd + cuBLAS (~95 launches);
                                    # needs fused_loss.  None: on for CUDA devices, off on the CPU
    fused_act: Optional[bool] = None   # rollout policy inference (normalise, MLP, tanh-normal sample, log-prob) as ONE kernel
                                    # (rr_policy_act) instead of ~20 torch launches per env step; needs 32-wide hidden policy layers
                                    # and <= 32 actions.  None: on for CUDA devices when the policy fits, off on the CPU
    fused_loss: bool = True         # loss + gradient w.r.t. the network outputs in two hand-written kernels (rr_ppo_loss)
    cuda_graph: bool = True         # replay the minibatch update (loss, backward, Adam) as one CUDA graph on CUDA devices
    graph_allreduce: bool = False   # several ranks: capture the NCCL all-reduce + Adam in the update graph too.  OFF: with torch 2.11 /
                                    # NCCL 2.28 the 2-rank capture hung on the B200 box (round 2); the eager all-reduce + Adam path is kept
    rollout_graph: bool = True      # replay the unroll (unroll_length x [policy MLP + sample + env step]) as one CUDA graph
    tf32: bool = True               # TF32 tensor-core matmuls, as XLA's default float32 dot precision on NVIDIA GPUs
    policy_hidden: tuple = (32, 32, 32, 32)
    value_hidden: tuple = (256, 256, 256, 256, 256)


def _mlp(sizes, out) -> nn.Sequential:
    layers, last = [], sizes[0]
    for h in sizes[1:]:
        layers += [nn.Linear(last, h), nn.SiLU()]
        last = h
    layers.append(nn.Linear(last, out))
    net = nn.Sequential(*layers)
    for mod in net:
        if isinstance(mod, nn.Linear):  # jax.nn.initializers.lecun_uniform, zero bias
            bound = math.sqrt(3.0 / mod.in_features)
            nn.init.uniform_(mod.weight, -bound, bound)
            nn.init.zeros_(mod.bias)
    return net


class RunningStats:
    """brax.training.acme.running_statistics: count / mean / summed_variance / std, updated with (optionally
    all-reduced) batch moments."""

    def __init__(self, size: int, device):
        self.count = torch.zeros((), device=device, dtype=torch.float64)
        self.mean = torch.zeros(size, device=device)
        self.summed_variance = torch.zeros(size, device=device)
        self.std = torch.ones(size, device=device)

    def update(self, batch: torch.Tensor, distributed: bool = False, std_min=1e-6, std_max=1e6) -> None:
        batch = batch.reshape(-1, batch.shape[-1])
        n = torch.tensor(float(batch.shape[0]), device=batch.device, dtype=torch.float64)
        diff_old = batch - self.mean
        s1 = diff_old.sum(0)
        if distributed:
            dist.all_reduce(n)
            dist.all_reduce(s1)
        count = self.count + n
        mean = self.mean + s1 / count.float()
        s2 = (diff_old * (batch - mean)).sum(0)
        if distributed:
            dist.all_reduce(s2)
        # in place: the captured rollout graph reads mean / std at fixed addresses
        self.summed_variance.add_(s2)
        self.count.copy_(count)
        self.mean.copy_(mean)
        self.std.copy_(torch.sqrt((self.summed_variance / count.float()).clamp_min(0.0)).clamp(std_min, std_max))

    def normalize(self, x: torch.Tensor) -> torch.Tensor:
        return (x - self.mean) / self.std

    def state_dict(self):
        return dict(count=self.count, mean=self.mean, summed_variance=self.summed_variance, std=self.std)

    def load_state_dict(self, d):
        for k in ("count", "mean", "summed_variance", "std"):
            getattr(self, k).copy_(torch.as_tensor(d[k], device=self.mean.device))


_LOG2 = math.log(2.0)


def tanh_normal_sample(logits: torch.Tensor, gen: Optional[torch.Generator] = None, eps: Optional[torch.Tensor] = None):
    """NormalTanhDistribution: returns (action, raw_action, log_prob).  `eps`: pre-drawn standard normal noise (the captured
    rollout graph draws it outside the graph)."""
    loc, scale = logits.chunk(2, dim=-1)
    scale = F.softplus(scale) + 1e-3
    if eps is None:
        eps = torch.randn(loc.shape, device=loc.device, generator=gen)
    raw = loc + scale * eps
    return torch.tanh(raw), raw, tanh_normal_log_prob(logits, raw)


def tanh_normal_params(logits: torch.Tensor):
    """(loc, scale, log scale) of NormalTanhDistribution; contiguous halves so that the element-wise kernels that follow are
    the vectorised ones (a strided `chunk` view sends every one of them down the generic strided path)."""
    a = logits.shape[-1] // 2
    loc, pre = logits[..., :a].contiguous(), logits[..., a:].contiguous()
    scale = F.softplus(pre) + 1e-3
    return loc, scale, torch.log(scale)


def _log_det_tanh(raw: torch.Tensor) -> torch.Tensor:
    return 2.0 * (_LOG2 - raw - F.softplus(-2.0 * raw))


def tanh_normal_log_prob(logits: torch.Tensor, raw: torch.Tensor, params=None) -> torch.Tensor:
    loc, scale, log_scale = params if params is not None else tanh_normal_params(logits)
    lp = -0.5 * ((raw - loc) / scale) ** 2 - log_scale - 0.5 * math.log(2 * math.pi)
    return (lp - _log_det_tanh(raw)).sum(-1)


def tanh_normal_entropy(logits: torch.Tensor, gen: Optional[torch.Generator] = None,
                        noise: Optional[torch.Tensor] = None, params=None) -> torch.Tensor:
    loc, scale, log_scale = params if params is not None else tanh_normal_params(logits)
    if noise is None:
        noise = torch.randn(loc.shape, device=loc.device, generator=gen)
    raw = loc + scale * noise
    return (0.5 + 0.5 * math.log(2 * math.pi) + log_scale + _log_det_tanh(raw)).sum(-1)


def compute_gae(L, truncation, termination, rewards, values, bootstrap, lambda_, discount):
    """ppo.losses.compute_gae on the rr_gae kernel; inputs time-major [T, B] contiguous CUDA tensors."""
    T, B = rewards.shape
    vs, adv = torch.empty_like(rewards), torch.empty_like(rewards)
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream) if rewards.is_cuda else None
    _lib.check(L, L.rr_gae(p(rewards.contiguous()), p(values.contiguous()), p(bootstrap.contiguous()), p(termination.contiguous()),
                           p(truncation.contiguous()), T, B, discount, lambda_, p(vs), p(adv), stream))
    return vs, adv


def launch_ppo_loss(L, cfg, logits, baseline, bootstrap, mb, noise, grad_logits=None, grad_baseline=None):
    """rr_ppo_loss on time-major logits [T, B, 2A], baseline [T, B], bootstrap [B]: returns (loss_partial [blocks, 3] -- column
    sums / (T B) are the policy, value and entropy terms --, d total / d logits, d total / d baseline)."""
    T, B = baseline.shape
    A = logits.shape[-1] // 2
    dev = logits.device
    c = lambda x: x.detach().contiguous()
    logits_c, baseline_c = c(logits), c(baseline)
    ba, bb = ctypes.c_int32(), ctypes.c_int32()
    _lib.check(L, L.rr_ppo_loss_blocks(T, B, ctypes.byref(ba), ctypes.byref(bb)))
    scratch = torch.empty(3 * T * B, device=dev)
    adv_partial = torch.empty(2 * ba.value, device=dev, dtype=torch.float64)
    loss_partial = torch.empty((bb.value, 3), device=dev)
    if grad_logits is None:
        grad_logits, grad_baseline = torch.empty_like(logits_c), torch.empty_like(baseline_c)
    keep = [logits_c, baseline_c, c(bootstrap), c(mb["raw_action"]), c(mb["log_prob"]), c(mb["reward"]), c(mb["discount"]),
            c(mb["truncation"]), c(noise)]
    a = _lib.RRPpoLossArgs()
    a.T, a.B, a.A = T, B, A
    for name, t in zip(("logits", "baseline", "bootstrap", "raw_action", "old_log_prob", "reward", "discount", "truncation",
                        "noise"), keep):
        assert t.dtype == torch.float32
        setattr(a, name, t.data_ptr())
    a.reward_scaling, a.discounting, a.gae_lambda = cfg.reward_scaling, cfg.discounting, cfg.gae_lambda
    a.clipping_epsilon, a.entropy_cost, a.normalize_advantage = cfg.clipping_epsilon, cfg.entropy_cost, int(cfg.normalize_advantage)
    a.scratch, a.adv_partial, a.loss_partial = scratch.data_ptr(), adv_partial.data_ptr(), loss_partial.data_ptr()
    a.grad_logits, a.grad_baseline = grad_logits.data_ptr(), grad_baseline.data_ptr()
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if logits.is_cuda else None
    _lib.check(L, L.rr_ppo_loss(ctypes.byref(a), stream))
    return loss_partial, grad_logits, grad_baseline


class _FusedPPOLoss(torch.autograd.Function):
    """total loss of one minibatch through rr_ppo_loss; backward hands the kernel's analytic d loss / d (logits, baseline)
    to autograd, which continues through the two MLPs.  Returns (total, policy_loss, v_loss, entropy_loss)."""

    @staticmethod
    def forward(ctx, logits, baseline, bootstrap, mb, noise, cfg, L):
        loss_partial, grad_logits, grad_baseline = launch_ppo_loss(L, cfg, logits, baseline, bootstrap, mb, noise)
        T, B = baseline.shape
        terms = loss_partial.sum(0) / float(T * B)
        policy_loss, v_loss, entropy_loss = terms[0], terms[1], -cfg.entropy_cost * terms[2]
        ctx.save_for_backward(grad_logits, grad_baseline)
        ctx.mark_non_differentiable(policy_loss, v_loss, entropy_loss)
        return policy_loss + v_loss + entropy_loss, policy_loss, v_loss, entropy_loss

    @staticmethod
    def backward(ctx, g_total, *_):
        grad_logits, grad_baseline = ctx.saved_tensors
        return g_total * grad_logits, g_total * grad_baseline, None, None, None, None, None


class _LazyLossMetrics(collections.abc.Mapping):
    """The loss terms of an update, formed from the loss kernel's per-block sums only when somebody reads them: inside the
    captured update they would be five more launches per minibatch for numbers that are read once per training step."""

    _KEYS = ("total_loss", "policy_loss", "v_loss", "entropy_loss")

    def __init__(self, loss_partial, n, entropy_cost):
        self._partial, self._n, self._ec = loss_partial, float(n), entropy_cost

    def _terms(self):
        t = self._partial.sum(0) / self._n
        policy_loss, v_loss, entropy_loss = t[0], t[1], -self._ec * t[2]
        return dict(total_loss=policy_loss + v_loss + entropy_loss, policy_loss=policy_loss, v_loss=v_loss, entropy_loss=entropy_loss)

    def __getitem__(self, k):
        return self._terms()[k]

    def __iter__(self):
        return iter(self._KEYS)

    def __len__(self):
        return len(self._KEYS)

    def items(self):
        return self._terms().items()


class PPO:
    def __init__(self, env: Rodent, cfg: PPOConfig):
        self.env, self.cfg = env, cfg
        self.device = env.device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        assert (cfg.batch_size * cfg.num_minibatches) % cfg.num_envs == 0, "batch_size * num_minibatches % num_envs"
        assert env.num_envs == cfg.num_envs
        torch.manual_seed(cfg.seed)  # identical initial parameters on every rank
        obs, act = env.observation_size, env.action_size
        # The first layers take the observation zero-padded to a multiple of 16 features (1263 -> 1264): with an odd K the
        # [5120 x 1263] x [1263 x 256] GEMMs fall to cuBLAS's unaligned kernels (3x slower).  The pad weights see only zeros:
        # they get no gradient and are dropped by export_brax_params.
        self._obs_dim, self._obs_pad = obs, (-obs) % 16 if self.device.type == "cuda" else 0
        self.policy = _mlp((obs + self._obs_pad,) + tuple(cfg.policy_hidden), 2 * act).to(self.device)
        self.value = _mlp((obs + self._obs_pad,) + tuple(cfg.value_hidden), 1).to(self.device)
        self.params = list(self.policy.parameters()) + list(self.value.parameters())
        self._use_graph = bool(cfg.cuda_graph) and self.device.type == "cuda"
        if self.device.type == "cuda" and cfg.tf32:
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        cuda = self.device.type == "cuda"
        # tensor-core learner (tc_learner.py): parameters, gradients and the Adam state live in flat buffers from the start (the
        # rollout graph captures the parameters' addresses); the GEMM launch lists are built at the first loss_and_grads() call
        self._use_tc = cuda if cfg.tc_learner is None else bool(cfg.tc_learner)
        assert not self._use_tc or cfg.fused_loss, "tc_learner needs fused_loss"
        self._tc = None
        if self._use_tc:
            from .tc_learner import FlatAdam, flatten_parameters
            self._flat_param, self._flat_g, self._flat_off = flatten_parameters(self.params, self.device)
            self.opt = FlatAdam(env._L, self._flat_param, self._flat_g, cfg.learning_rate, eps=1e-8)
        else:
            self.opt = torch.optim.Adam(self.params, lr=cfg.learning_rate, eps=1e-8, capturable=self._use_graph,
                                        fused=True if cuda else None)  # one multi-tensor kernel instead of ~15 foreach launches
        self._use_rollout_graph = bool(cfg.rollout_graph) and self.device.type == "cuda"
        self._rg = None             # captured unroll (policy + sample + env step) x unroll_length
        self._graph = None          # (gather + fwd + bwd [+ Adam]) graph, static batch / minibatch buffers, static metrics
        self._static = None
        self._graph_warm = 0
        self.normalizer = RunningStats(obs, self.device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(cfg.seed * 1000 + 17 + self.rank)
        self.env_steps = 0
        self._flat_grad = None
        self._use_fused_act = self._fused_act_ok()
        self._batch_static = None
        self._gather_items = None

    # ---- acting -----------------------------------------------------------------------------------------------------
    def _norm(self, obs):
        obs = self.normalizer.normalize(obs) if self.cfg.normalize_observations else obs
        return F.pad(obs, (0, self._obs_pad)) if self._obs_pad else obs

    def _norm_mb(self, obs):
        """Observations of a minibatch: already normalised when they come out of training_step's batch."""
        return obs if getattr(self, "_batch_is_normalized", False) else self._norm(obs)

    def _fused_act_ok(self) -> bool:
        lin = [m for m in self.policy if isinstance(m, nn.Linear)]
        fits = (2 <= len(lin) <= _lib.RR_POLICY_MAX_LAYERS and all(l.out_features == 32 for l in lin[:-1]) and
                self.env.action_size <= 32 and lin[-1].out_features == 2 * self.env.action_size)
        want = (self.device.type == "cuda") if self.cfg.fused_act is None else bool(self.cfg.fused_act)
        if self.cfg.fused_act and not fits:
            raise ValueError("fused_act needs 32-wide hidden policy layers and at most 32 actions")
        return want and fits

    def _act_fused(self, obs, deterministic, eps):
        """rr_policy_act: one kernel from the raw observation to (action, raw action, log-prob)."""
        B, A = obs.shape[0], self.env.action_size
        obs = obs.contiguous()
        if not deterministic and eps is None:
            eps = torch.randn((B, A), device=obs.device, generator=self.gen)
        action, raw, lp = (torch.empty((B, A), device=obs.device), torch.empty((B, A), device=obs.device),
                           torch.empty(B, device=obs.device))
        lin = [m for m in self.policy if isinstance(m, nn.Linear)]
        a = _lib.RRPolicyArgs()
        a.obs = obs.data_ptr()
        if self.cfg.normalize_observations:
            a.mean, a.std = self.normalizer.mean.data_ptr(), self.normalizer.std.data_ptr()
        for i, l in enumerate(lin):
            a.w[i], a.b[i] = l.weight.data_ptr(), l.bias.data_ptr()
        a.eps = None if deterministic else eps.contiguous().data_ptr()
        a.action, a.raw_action, a.log_prob = action.data_ptr(), raw.data_ptr(), lp.data_ptr()
        a.B, a.obs_dim, a.in0, a.nlayers, a.A = B, self._obs_dim, lin[0].in_features, len(lin), A
        stream = ctypes.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream) if obs.is_cuda else None
        _lib.check(self.env._L, self.env._L.rr_policy_act(ctypes.byref(a), stream))
        if deterministic:
            lp.zero_()
        return action, raw, lp

    @torch.no_grad()
    def act(self, obs, deterministic=False, eps=None):
        if self._use_fused_act:
            return self._act_fused(obs, deterministic, eps)
        logits = self.policy(self._norm(obs))
        if deterministic:
            loc = logits.chunk(2, dim=-1)[0]
            return torch.tanh(loc), loc, torch.zeros(obs.shape[0], device=obs.device)
        return tanh_normal_sample(logits, self.gen, eps)

    @torch.no_grad()
    def unroll(self, state: State):
        """acting.generate_unroll: `unroll_length` policy + env steps; returns (state, transitions time-major).  On CUDA the
        whole unroll is captured once as a CUDA graph and replayed (cfg.rollout_graph)."""
        if self._use_rollout_graph:
            return self._unroll_graphed(state)
        return self._unroll_eager(state, None)

    # fields of the carried State that the captured unroll reads at fixed addresses
    _PS_FIELDS = ("qpos", "qvel", "act", "qacc_warmstart", "time")

    def _state_tensors(self, state: State):
        ps = state.pipeline_state
        out = [getattr(ps, k) for k in self._PS_FIELDS] + [state.obs, state.reward, state.done, state.info["cur_frame"]]
        out += [state.info[k] for k in ("steps", "truncation") if k in state.info]
        return out

    @torch.no_grad()
    def _unroll_graphed(self, state: State):
        """The unroll as ONE CUDA graph: policy MLP + tanh-normal sample + rr_step_kernel, unroll_length times.  The carried
        State lives in static tensors (the graph ends by copying the final state back into them), the exploration noise is
        drawn outside the graph from the agent's generator, and the transitions are copied out of the graph's memory after
        every replay (they are overwritten by the next one).  Same arithmetic as the eager unroll (tests/test_ppo.py)."""
        T, B, A = self.cfg.unroll_length, self.env.num_envs, self.env.action_size
        if self._rg is None:
            # static carried state = clones of the first state handed in (the auto-reset cache first_* is shared, read-only)
            import copy
            ps = copy.copy(state.pipeline_state)
            for k in self._PS_FIELDS:
                setattr(ps, k, getattr(ps, k).clone())
            info = dict(state.info)
            for k in ("cur_frame", "steps", "truncation"):
                if k in info:
                    info[k] = info[k].clone()
            self._rg_state = State(ps, state.obs.clone(), state.reward.clone(), state.done.clone(), dict(state.metrics), info)
            self._rg_eps = torch.empty((T, B, A), device=self.device)
            self._rg_eps.normal_(generator=self.gen)
            side = torch.cuda.Stream(self.device)   # warm-up off the capture stream (allocator, cuBLAS workspaces)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._unroll_eager(self._rg_state, self._rg_eps)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                new_state, data = self._unroll_eager(self._rg_state, self._rg_eps)
                for dst, src in zip(self._state_tensors(self._rg_state), self._state_tensors(new_state)):
                    dst.copy_(src)
            self._rg, self._rg_data = g, data
        elif state is not self._rg_state:  # a state from elsewhere (e.g. a fresh reset): load it into the static buffers
            for dst, src in zip(self._state_tensors(self._rg_state), self._state_tensors(state)):
                dst.copy_(src)
        self._rg_eps.normal_(generator=self.gen)
        self._rg.replay()
        data = {k: v.clone() for k, v in self._rg_data.items()}
        return self._rg_state, data

    @torch.no_grad()
    def _unroll_eager(self, state: State, eps_all):
        T = self.cfg.unroll_length
        obs, raw, logp, rew, disc, trunc = [], [], [], [], [], []
        for t in range(T):
            action, raw_a, lp = self.act(state.obs, eps=None if eps_all is None else eps_all[t])
            obs.append(state.obs)
            state = self.env.step(state, action)
            raw.append(raw_a); logp.append(lp); rew.append(state.reward)
            disc.append(1.0 - state.done)
            trunc.append(state.info["truncation"])
        data = dict(observation=torch.stack(obs), raw_action=torch.stack(raw), log_prob=torch.stack(logp),
                    reward=torch.stack(rew), discount=torch.stack(disc), truncation=torch.stack(trunc),
                    next_observation_last=state.obs)
        return state, data

    # ---- learning ---------------------------------------------------------------------------------------------------
    def loss(self, mb: Dict[str, torch.Tensor]):
        cfg = self.cfg
        obs = self._norm_mb(mb["observation"])                   # [T, b, obs]
        logits = self.policy(obs)
        baseline = self.value(obs).squeeze(-1)
        if cfg.fused_loss:
            with torch.no_grad():
                bootstrap = self.value(self._norm_mb(mb["next_observation_last"])).squeeze(-1)
                noise = mb.get("entropy_noise")
                if noise is None:
                    noise = torch.randn(mb["raw_action"].shape, device=obs.device, generator=self.gen)
            total, policy_loss, v_loss, entropy_loss = _FusedPPOLoss.apply(logits, baseline, bootstrap, mb, noise, cfg, self.env._L)
            return total, dict(total_loss=total.detach(), policy_loss=policy_loss, v_loss=v_loss, entropy_loss=entropy_loss)
        with torch.no_grad():
            bootstrap = self.value(self._norm_mb(mb["next_observation_last"])).squeeze(-1)
            rewards = mb["reward"] * cfg.reward_scaling
            truncation = mb["truncation"]
            termination = (1 - mb["discount"]) * (1 - truncation)
            vs, adv = compute_gae(self.env._L, truncation, termination, rewards, baseline.detach(), bootstrap, cfg.gae_lambda,
                                  cfg.discounting)
            if cfg.normalize_advantage:
                adv = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
        dist_params = tanh_normal_params(logits)
        target_lp = tanh_normal_log_prob(logits, mb["raw_action"], dist_params)
        rho = torch.exp(target_lp - mb["log_prob"])
        s1, s2 = rho * adv, rho.clamp(1 - cfg.clipping_epsilon, 1 + cfg.clipping_epsilon) * adv
        policy_loss = -torch.min(s1, s2).mean()
        v_loss = ((vs - baseline) ** 2).mean() * 0.5 * 0.5
        entropy = tanh_normal_entropy(logits, self.gen, mb.get("entropy_noise"), dist_params).mean()
        entropy_loss = -cfg.entropy_cost * entropy
        total = policy_loss + v_loss + entropy_loss
        return total, dict(total_loss=total.detach(), policy_loss=policy_loss.detach(), v_loss=v_loss.detach(),
                           entropy_loss=entropy_loss.detach())

    def loss_and_grads(self, mb: Dict[str, torch.Tensor], zero_grad_to_none: bool = False, defer_grad_sum: bool = False):
        """Loss of one minibatch and its gradient in every parameter's .grad.  Autograd through cuBLAS, or (cfg.tc_learner) the
        hand-written forward / backward on the tcgen05 GEMM kernel: 12 grouped launches + the two loss kernels.  `defer_grad_sum`:
        the caller steps the optimizer right after (the .grad tensors are complete only after that step)."""
        if not self._use_tc:
            total, metrics = self.loss(mb)
            self.opt.zero_grad(set_to_none=zero_grad_to_none)
            total.backward()
            return metrics
        cfg = self.cfg
        obs, nxt = self._norm_mb(mb["observation"]), self._norm_mb(mb["next_observation_last"])
        T, b = obs.shape[0], obs.shape[1]
        if self._tc is None or self._tc.x.shape[0] != T * b:
            from .tc_learner import TcLearner
            self._tc = TcLearner(self.env._L, self.policy, self.value, T * b, b, self.device, self._flat_g, self._flat_off)
        tc = self._tc
        if obs.data_ptr() != tc.x.data_ptr():
            tc.x.copy_(obs.reshape(T * b, -1))      # tc.x is a pitched view (the buffer carries a column of ones)
        if nxt.data_ptr() != tc.xb.data_ptr():
            tc.xb.copy_(nxt.reshape(b, -1))
        tc.forward()
        noise = mb.get("entropy_noise")
        if noise is None:
            noise = torch.randn(mb["raw_action"].shape, device=obs.device, generator=self.gen)
        loss_partial, _, _ = launch_ppo_loss(self.env._L, cfg, tc.logits.view(T, b, -1), tc.baseline.view(T, b), tc.bootstrap.view(b),
                                             mb, noise, tc.grad_logits, tc.grad_baseline)
        # one rank: the optimizer's kernel adds the split weight gradients' partial sums itself (the all-reduce needs them summed)
        tc.backward(defer_sum_to=self.opt if (defer_grad_sum and self.world == 1) else None)
        return _LazyLossMetrics(loss_partial, T * b, cfg.entropy_cost)

    def _allreduce_grads(self):
        """lax.pmean(grads): one flat fp32 bucket (2.53 MB for the rodent networks) per minibatch over NCCL."""
        if self.world == 1:
            return
        if self._use_tc and self._tc is not None:   # the gradients already live in one flat buffer
            dist.all_reduce(self._tc.flat_grad)
            self._tc.flat_grad.div_(self.world)
            return
        grads = [p.grad for p in self.params]
        if self._flat_grad is None:
            self._flat_grad = torch.empty(sum(g.numel() for g in grads), device=self.device)
        torch.cat([g.reshape(-1) for g in grads], out=self._flat_grad)
        dist.all_reduce(self._flat_grad)
        self._flat_grad.div_(self.world)
        off = 0
        for g in grads:
            g.copy_(self._flat_grad[off:off + g.numel()].view_as(g))
            off += g.numel()

    def _update_graphed(self, data: Dict[str, torch.Tensor], idx: torch.Tensor):
        """One minibatch update with the loss / backward / Adam kernels replayed as a CUDA graph (the eager update is
        ~150 small launches and host-bound at ~3 ms; the GPU work is ~0.3 ms).  The minibatch is gathered into static
        buffers, the entropy noise is drawn outside the graph from the agent's generator.  With several ranks the NCCL
        all-reduce of the flat gradient bucket and Adam are captured in the same graph (cfg.graph_allreduce; off: the graph
        ends after backward and they run eagerly)."""
        cfg = self.cfg
        if self._graph is None and self._static is None:
            self._static = {k: torch.empty((v.shape[0], cfg.batch_size) + tuple(v.shape[2:]) if k != "next_observation_last"
                                           else (cfg.batch_size,) + tuple(v.shape[1:]), device=self.device, dtype=v.dtype)
                            for k, v in data.items()}
            self._static["entropy_noise"] = torch.empty((cfg.unroll_length, cfg.batch_size, self.env.action_size), device=self.device)
            if self._use_tc:  # gather the minibatch straight into the tensor-core learner's input buffers
                from .tc_learner import TcLearner
                T, b = data["observation"].shape[0], cfg.batch_size
                self._tc = TcLearner(self.env._L, self.policy, self.value, T * b, b, self.device, self._flat_g, self._flat_off)
                self._static["observation"] = self._tc.x.unflatten(0, (T, b))   # pitched view: rows of k0 floats, pitch k0 + 4
                self._static["next_observation_last"] = self._tc.xb
            # the batch the minibatches are gathered from and the minibatch's indices are static too, so that the gathers are
            # part of the graph (seven index_select launches per update were most of the host time of an update)
            self._batch_static = dict(data)  # adopt this batch's tensors; later batches are assembled in them (training_step)
            self._idx_static = torch.zeros(cfg.batch_size, dtype=idx.dtype, device=self.device)
        st, batch = self._static, self._batch_static
        for k, v in data.items():
            if v.data_ptr() != batch[k].data_ptr():
                batch[k].copy_(v)
        self._idx_static.copy_(idx)
        st["entropy_noise"].normal_(generator=self.gen)

        def gather():
            if self._gather_items is None:  # one launch for all seven tensors (rr_gather_rows)
                keys = list(batch)
                assert all(batch[k].dtype == torch.float32 and batch[k].is_contiguous() for k in keys)
                items = (_lib.RRGatherItem * len(keys))()
                for it, k in zip(items, keys):
                    v = batch[k]
                    lead = 2 if k != "next_observation_last" else 1      # [T, N, ...] or [N, ...]
                    it.src, it.dst = v.data_ptr(), st[k].data_ptr()
                    it.outer = v.shape[0] if lead == 2 else 1
                    it.src_rows = v.shape[lead - 1]
                    it.inner = int(math.prod(v.shape[lead:]))
                    d = st[k]                                                 # contiguous, or rows with a pitch (2-D inner only)
                    it.dst_pitch = it.inner if d.is_contiguous() else d.stride(lead - 1)
                    assert d.is_contiguous() or (d.dim() == lead + 1 and d.stride(-1) == 1 and
                                                 (lead == 1 or d.stride(0) == d.shape[1] * d.stride(1)))
                self._gather_items = items
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.env._L, self.env._L.rr_gather_rows(self._gather_items, len(self._gather_items),
                                                               ctypes.c_void_p(self._idx_static.data_ptr()), cfg.batch_size, stream))

        if self._graph is None and self._graph_warm < 3:
            # eager warm-up updates on a side stream (allocator / cuBLAS workspaces / Adam state), as torch's capture recipe
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                gather()
                metrics = self.loss_and_grads(st, zero_grad_to_none=True, defer_grad_sum=True)
                self._allreduce_grads()
                self.opt.step()
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph_warm += 1
            return metrics
        if self._graph is None:
            g = torch.cuda.CUDAGraph()
            if not self._use_tc:
                self.opt.zero_grad(set_to_none=True)
            in_graph = self.world == 1 or self.cfg.graph_allreduce
            with torch.cuda.graph(g):
                gather()
                metrics = self.loss_and_grads(st, zero_grad_to_none=True, defer_grad_sum=True)
                if in_graph:
                    self._allreduce_grads()  # NCCL all-reduce of the flat bucket, captured with the rest (no-op on one rank)
                    self.opt.step()
            self._graph, self._graph_metrics, self._graph_full = g, metrics, in_graph
        self._graph.replay()
        if not self._graph_full:
            self._allreduce_grads()
            self.opt.step()
        return self._graph_metrics

    def training_step(self, state: State):
        cfg = self.cfg
        n_unroll = cfg.batch_size * cfg.num_minibatches // cfg.num_envs
        chunks = []
        for _ in range(n_unroll):
            state, data = self.unroll(state)
            chunks.append(data)
        # [T, n_unroll * num_envs, ...]: the "batch" axis that brax shuffles is (unroll, env)
        obs_keys = ("observation", "next_observation_last")
        out = getattr(self, "_batch_out", None)   # the update graph's static batch buffers once it exists
        data = {k: torch.cat([c[k] for c in chunks], dim=1 if k != "next_observation_last" else 0,
                             out=out[k] if out is not None and k not in obs_keys else None) for k in chunks[0]}
        if cfg.normalize_observations:
            self.normalizer.update(data["observation"], distributed=self.world > 1)
        # the statistics are fixed for all epochs of this batch: normalise (and pad) it once instead of once per minibatch
        for k in obs_keys:
            normed = self._norm(data[k])
            if out is not None:
                out[k].copy_(normed)
                normed = out[k]
            data[k] = normed
        self._batch_is_normalized = True
        nb = data["reward"].shape[1]
        metrics = {}
        for _ in range(cfg.num_updates_per_batch):
            perm = torch.randperm(nb, device=self.device, generator=self.gen)
            for i in range(cfg.num_minibatches):
                idx = perm[i * cfg.batch_size:(i + 1) * cfg.batch_size]
                if self._use_graph:
                    metrics = self._update_graphed(data, idx)
                    continue
                mb = {k: (v[:, idx] if k != "next_observation_last" else v[idx]) for k, v in data.items()}
                metrics = self.loss_and_grads(mb, defer_grad_sum=True)
                self._allreduce_grads()
                self.opt.step()
        if self._use_graph:
            metrics = {k: v.clone() for k, v in metrics.items()}
            if self._batch_static is not None:  # from now on the next batches are assembled in the graph's static buffers
                self._batch_out = self._batch_static
        self._batch_is_normalized = False
        self.env_steps += n_unroll * cfg.unroll_length * cfg.num_envs * self.world
        return state, metrics

    # ---- evaluation (acting.Evaluator) ---------------------------------------------------------------------------------
    @torch.no_grad()
    def evaluate(self, eval_env: Rodent, seed: int = 1) -> Dict[str, float]:
        state = eval_env.reset(seed)
        ret = torch.zeros(eval_env.num_envs, device=self.device)
        alive = torch.ones_like(ret)
        length = torch.zeros_like(ret)
        sums = {k: torch.zeros_like(ret) for k in ("pos_reward", "reward_quadctrl", "reward_alive")}
        for _ in range(self.cfg.episode_length):
            action, _, _ = self.act(state.obs, deterministic=self.cfg.deterministic_eval)
            state = eval_env.step(state, action)
            ret += state.reward * alive
            length += alive
            for k in sums:
                sums[k] += state.metrics[k] * alive
            alive = alive * (1 - state.done)
        out = {"eval/episode_reward": ret.mean(), "eval/avg_episode_length": length.mean()}
        out.update({f"eval/episode_{k}": v.mean() for k, v in sums.items()})
        if self.world > 1:
            for v in out.values():
                dist.all_reduce(v)
                v.div_(self.world)
        return {k: float(v) for k, v in out.items()}

    # ---- checkpointing (brax model.save_params saves (normalizer, policy); we also keep value + optimiser) ---------------
    def state_dict(self):
        return dict(policy=self.policy.state_dict(), value=self.value.state_dict(), normalizer=self.normalizer.state_dict(),
                    optimizer=self.opt.state_dict(), env_steps=self.env_steps)

    def load_state_dict(self, d):
        self.policy.load_state_dict(d["policy"]); self.value.load_state_dict(d["value"])
        self.normalizer.load_state_dict(d["normalizer"]); self.opt.load_state_dict(d["optimizer"])
        self.env_steps = d["env_steps"]
        self._graph, self._graph_warm = None, 0  # the captured Adam kernels point at the replaced optimiser state: recapture

    def export_brax_params(self):
        """(normalizer, policy) in the flax naming brax pickles (`hidden_i` / kernel [in, out] / bias), as numpy."""
        def mlp(net):
            lin = [m for m in net if isinstance(m, nn.Linear)]
            return {"params": {f"hidden_{i}": {"kernel": l.weight.detach().t()[:self._obs_dim if i == 0 else None].cpu().numpy(),
                                               "bias": l.bias.detach().cpu().numpy()} for i, l in enumerate(lin)}}
        n = self.normalizer
        norm = dict(count=float(n.count), mean=n.mean.cpu().numpy(), summed_variance=n.summed_variance.cpu().numpy(),
                    std=n.std.cpu().numpy())
        return norm, mlp(self.policy)


    def import_brax_params(self, params) -> None:
        """Load `(normalizer, policy)` as pickled by brax `model.save_params` (brax_rodent_run_ppo.py:204-206,
        render_rollout.ipynb:117-118) or by `export_brax_params`: the normaliser may be a dict or any object with
        count / mean / std (/ summed_variance); the policy a flax param tree {"params": {"hidden_i": {kernel, bias}}}."""
        norm, pol = params[0], params[1]
        get = (lambda k: norm.get(k)) if isinstance(norm, dict) else (lambda k: getattr(norm, k, None))
        n = self.normalizer
        as_t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32, device=self.device)
        # in place: the captured rollout graph (and rr_policy_act's argument block) read the normaliser at fixed addresses
        n.count.copy_(torch.as_tensor(float(np.asarray(get("count"))), dtype=torch.float64, device=self.device))
        n.mean.copy_(as_t(get("mean")))
        n.std.copy_(as_t(get("std")))
        sv = get("summed_variance")
        n.summed_variance.copy_(as_t(sv) if sv is not None else (n.std ** 2) * float(n.count))
        layers = pol["params"] if "params" in pol else pol
        lin = [m for m in self.policy if isinstance(m, nn.Linear)]
        if len(layers) != len(lin):
            raise ValueError(f"policy has {len(lin)} layers, the parameters {len(layers)}")
        with torch.no_grad():
            for i, l in enumerate(lin):
                w, b = as_t(layers[f"hidden_{i}"]["kernel"]).t(), as_t(layers[f"hidden_{i}"]["bias"])
                if i == 0 and self._obs_pad:
                    w = F.pad(w, (0, self._obs_pad))
                if w.shape != l.weight.shape:
                    raise ValueError(f"hidden_{i}: kernel {tuple(w.t().shape)} does not fit {tuple(l.weight.t().shape)}")
                l.weight.copy_(w); l.bias.copy_(b)
        self._graph, self._graph_warm = None, 0


def policy_from_brax_params(params, device, deterministic: bool = False, generator: Optional[torch.Generator] = None):
    """brax `make_policy(params, deterministic)` (ppo_networks.make_inference_fn) for a `(normalizer, policy)` parameter tuple
    as pickled by brax model.save_params / returned by PPO.export_brax_params: returns policy(obs, rng=None) -> (action, extras)
    with the observation normalised by (mean, std) and a swish MLP in flax naming.  Independent of any PPO instance, so a
    callback can evaluate a checkpoint while training goes on (brax_rodent_run_ppo.py:135-151)."""
    norm, pol = params[0], params[1]
    get = (lambda k: norm.get(k)) if isinstance(norm, dict) else (lambda k: getattr(norm, k, None))
    as_t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32, device=device)
    mean, std = as_t(get("mean")), as_t(get("std"))
    layers = pol["params"] if "params" in pol else pol
    ws = [(as_t(layers[f"hidden_{i}"]["kernel"]), as_t(layers[f"hidden_{i}"]["bias"])) for i in range(len(layers))]

    @torch.no_grad()
    def policy(obs, rng=None):
        h = (obs.to(device, torch.float32) - mean) / std
        for i, (w, b) in enumerate(ws):
            h = h @ w + b
            if i + 1 < len(ws):
                h = F.silu(h)
        if deterministic:
            loc = h.chunk(2, dim=-1)[0]
            return torch.tanh(loc), {"log_prob": torch.zeros(obs.shape[0], device=device), "raw_action": loc}
        gen = rng if isinstance(rng, torch.Generator) else generator
        action, raw, lp = tanh_normal_sample(h, gen)
        return action, {"log_prob": lp, "raw_action": raw}

    return policy


def train(environment: Rodent, cfg: PPOConfig, progress_fn: Callable[[int, Dict], None] = lambda *a: None,
          policy_params_fn: Callable = lambda *a: None, eval_env: Optional[Rodent] = None):
    """ppo.train(environment=env, progress_fn=..., policy_params_fn=...) (brax_rodent_run_ppo.py:200-202).
    Returns (make_inference_fn, params, metrics) like brax.

    Schedule as brax.training.agents.ppo.train: `num_evals_after_init = max(num_evals - 1, 1)` epochs of
    `ceil(num_timesteps / (num_evals_after_init * env_steps_per_training_step))` training steps, each epoch followed by
    evaluation, progress_fn(num_steps, metrics) and policy_params_fn(num_steps, make_policy, params); with num_evals > 1 the
    untrained policy is evaluated first (progress_fn(0, ...)).  `make_policy(params, deterministic=False)` is brax's
    make_inference_fn: called with the `params` handed to the callback (or any exported tuple) it builds a stand-alone
    policy; called without arguments it acts with the live agent."""
    agent = PPO(environment.wrap_for_training(cfg.episode_length), cfg)
    steps_per_train = cfg.batch_size * cfg.num_minibatches * cfg.unroll_length * agent.world
    evals_after_init = max(cfg.num_evals - 1, 1)
    per_epoch = max(1, -(-cfg.num_timesteps // (evals_after_init * steps_per_train)))
    state = environment.reset(cfg.seed + agent.rank)
    metrics: Dict[str, float] = {}

    def make_inference_fn(params=None, deterministic=False):
        if params is not None and not isinstance(params, PPO):
            return policy_from_brax_params(params, agent.device, deterministic, agent.gen)

        def policy(obs, rng=None):
            action, raw, lp = agent.act(obs, deterministic)
            return action, {"log_prob": lp, "raw_action": raw}
        return policy

    if eval_env is not None and cfg.num_evals > 1:  # brax evaluates the untrained policy first (`num_evals` counts it)
        m0 = agent.evaluate(eval_env)
        if agent.rank == 0:
            progress_fn(0, m0)
    t0 = time.time()
    for epoch in range(evals_after_init):
        for _ in range(per_epoch):
            state, m = agent.training_step(state)
        if agent.device.type == "cuda":
            torch.cuda.synchronize(agent.device)
        metrics = {f"training/{k}": float(v) for k, v in m.items()}
        metrics["training/sps"] = agent.env_steps / (time.time() - t0)
        metrics["training/walltime"] = time.time() - t0
        if eval_env is not None:
            metrics.update(agent.evaluate(eval_env))
        if agent.rank == 0:
            progress_fn(agent.env_steps, metrics)
            policy_params_fn(agent.env_steps, make_inference_fn, agent.export_brax_params())

    return make_inference_fn, agent, metrics
