"""PPO learner / rollout loop host logic (brax ppo.train restatement) on the emulator backend, incl. the world_size-2
gloo path that stands in for the NCCL gradient / normaliser all-reduce."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import EMU_LIB, load_asset, synthetic_track

TINY = dict(num_envs=2, batch_size=2, num_minibatches=2, unroll_length=2, num_updates_per_batch=2, episode_length=5,
            num_timesteps=16, policy_hidden=(8, 8), value_hidden=(16, 16))


def tiny_env(emu_lib, num_envs=2):
    from brax_rodent_run_b200.env import Rodent
    return Rodent(synthetic_track(), num_envs=num_envs, device="cpu", model=load_asset("rodent_0"), iterations=1, ls_iterations=1,
                  n_frames=1, _lib_path=emu_lib)


def test_tanh_normal_matches_torch_distributions():
    from brax_rodent_run_b200 import ppo
    torch.manual_seed(0)
    logits = torch.randn(7, 12)
    act, raw, lp = ppo.tanh_normal_sample(logits)
    loc, scale = logits.chunk(2, -1)
    scale = torch.nn.functional.softplus(scale) + 1e-3
    base = torch.distributions.Normal(loc, scale)
    want = (base.log_prob(raw) - torch.log(1 - torch.tanh(raw) ** 2 + 1e-12)).sum(-1)
    assert torch.allclose(lp, want, atol=1e-4)
    assert torch.allclose(act, torch.tanh(raw))


def test_running_stats_matches_numpy():
    from brax_rodent_run_b200.ppo import RunningStats
    rs = RunningStats(5, "cpu")
    rng = np.random.default_rng(0)
    xs = [rng.normal(2, 3, (11, 4, 5)).astype(np.float32) for _ in range(3)]
    for x in xs:
        rs.update(torch.tensor(x))
    allx = np.concatenate([x.reshape(-1, 5) for x in xs])
    assert np.allclose(rs.mean.numpy(), allx.mean(0), atol=1e-4)
    assert np.allclose(rs.std.numpy(), allx.std(0), atol=1e-3)


def test_training_step_runs_and_learns_shapes(emu_lib):
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    cfg = PPOConfig(**TINY)
    env = tiny_env(emu_lib).wrap_for_training(cfg.episode_length)
    agent = PPO(env, cfg)
    before = [p.detach().clone() for p in agent.params]
    state = env.reset(0)
    state, metrics = agent.training_step(state)
    assert agent.env_steps == cfg.batch_size * cfg.num_minibatches * cfg.unroll_length
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.params))
    assert float(agent.normalizer.count) == cfg.batch_size * cfg.num_minibatches * cfg.unroll_length
    norm, pol = agent.export_brax_params()
    assert pol["params"]["hidden_0"]["kernel"].shape == (env.observation_size, 8)
    sd = agent.state_dict()
    agent.load_state_dict(sd)
    # brax (normalizer, policy) parameter exchange: export -> pickle -> import into a fresh agent reproduces the policy
    import pickle
    other = PPO(env, PPOConfig(**dict(TINY, seed=7)))
    other.import_brax_params(pickle.loads(pickle.dumps(agent.export_brax_params())))
    a1, _, _ = agent.act(state.obs, deterministic=True)
    a2, _, _ = other.act(state.obs, deterministic=True)
    assert torch.allclose(a1, a2, atol=1e-6)
    # 8 samples make a degenerate normaliser (std clipped at 1e-6, as in brax); the loss itself is checked without it
    cfg2 = PPOConfig(**dict(TINY, normalize_observations=False))
    agent2 = PPO(env, cfg2)
    state, metrics = agent2.training_step(state)
    assert all(math.isfinite(float(v)) for v in metrics.values()), metrics


def _rank_main(rank, world, port, emu_lib, out, tc=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    cfg = PPOConfig(**dict(TINY, tc_learner=tc))
    env = tiny_env(emu_lib).wrap_for_training(cfg.episode_length)
    agent = PPO(env, cfg)
    state = env.reset(100 + rank)  # per-rank env shard: different seeds
    # the normaliser is exercised (and all-reduced) but not applied: 16 samples give a degenerate std (see above)
    agent.cfg = PPOConfig(**dict(TINY, normalize_observations=False, tc_learner=tc))
    agent.normalizer.update(state.obs, distributed=True)
    state, _ = agent.training_step(state)
    flat = torch.cat([p.detach().reshape(-1) for p in agent.params])
    assert torch.isfinite(flat).all()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    means = [torch.zeros_like(agent.normalizer.mean) for _ in range(world)]
    dist.all_gather(means, agent.normalizer.mean)
    if rank == 0:
        out.put((bool(torch.equal(gathered[0], gathered[1])), bool(torch.equal(means[0], means[1])),
                 float(agent.normalizer.count), agent.env_steps))
    dist.destroy_process_group()


@pytest.mark.parametrize("tc", [False, True], ids=["autograd", "tc_learner"])
def test_two_rank_gloo_training_step(emu_lib, tc):
    """Ranks hold different env shards; after a training step the parameters and the normaliser must be identical on both
    ranks (gradient all-reduce mean, normaliser moment all-reduce), and the counters are global.  With the tensor-core learner's
    launch lists (emulator backend) the all-reduce runs on the flat gradient buffer itself."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + (250 if tc else 0)) % 500
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, emu_lib, q, tc)) for r in range(2)]
    for p in procs:
        p.start()
    same_params, same_norm, count, env_steps = q.get(timeout=600)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg_steps = TINY["batch_size"] * TINY["num_minibatches"] * TINY["unroll_length"]
    assert same_params and same_norm
    assert count == 2 * TINY["num_envs"] and env_steps == 2 * cfg_steps


@pytest.mark.gpu
def test_graphed_update_matches_eager_update():
    """The CUDA-graph replay of the minibatch update (loss + backward + Adam) must report the same losses and apply the same
    parameter update as the eager code on the same minibatch, entropy noise, parameters and Adam state."""
    from brax_rodent_run_b200.env import Rodent
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    cfg = PPOConfig(num_envs=64, batch_size=16, num_minibatches=8, unroll_length=4, num_updates_per_batch=2, episode_length=50,
                    num_timesteps=1, policy_hidden=(32, 32), value_hidden=(64, 64), tf32=False, cuda_graph=True, tc_learner=False)
    env = Rodent(synthetic_track(), num_envs=64, device="cuda:0", model=load_asset("rodent_0"), iterations=4, ls_iterations=4,
                 terminate_when_unhealthy=False).wrap_for_training(cfg.episode_length)
    agent = PPO(env, cfg)
    state = env.reset(0)
    state, metrics = agent.training_step(state)          # warm-up updates, capture, replays
    assert agent._graph is not None
    assert all(math.isfinite(float(v)) for v in metrics.values())
    chunks = []
    for _ in range(cfg.batch_size * cfg.num_minibatches // cfg.num_envs):
        state, d = agent.unroll(state)
        chunks.append(d)
    data = {k: torch.cat([c[k] for c in chunks], dim=1 if k != "next_observation_last" else 0) for k in chunks[0]}
    # as training_step: the batch is normalised once, the captured graph and the eager loss then take it as it is
    data["observation"] = agent._norm(data["observation"])
    data["next_observation_last"] = agent._norm(data["next_observation_last"])
    agent._batch_is_normalized = True
    idx = torch.arange(3, 3 + cfg.batch_size, device="cuda:0")
    opt_tensors = [t for st in agent.opt.state.values() for t in st.values() if torch.is_tensor(t)]
    saved = [t.clone() for t in agent.params + opt_tensors]
    gen_state = agent.gen.get_state()

    def restore():
        with torch.no_grad():
            for t, v in zip(agent.params + opt_tensors, saved):
                t.copy_(v)
        agent.gen.set_state(gen_state)

    m_graph = {k: float(v) for k, v in agent._update_graphed(data, idx).items()}
    d_graph = torch.cat([(p.detach() - v).reshape(-1) for p, v in zip(agent.params, saved)]).cpu()
    restore()
    st = agent._static                                    # still holds the gathered minibatch
    st["entropy_noise"].normal_(generator=agent.gen)      # same generator state -> same noise
    total, m = agent.loss(st)
    agent.opt.zero_grad(set_to_none=False)
    total.backward()
    agent.opt.step()
    m_eager = {k: float(v) for k, v in m.items()}
    d_eager = torch.cat([(p.detach() - v).reshape(-1) for p, v in zip(agent.params, saved)]).cpu()
    for k in m_eager:
        assert abs(m_graph[k] - m_eager[k]) <= 1e-4 * max(1.0, abs(m_eager[k])), (k, m_graph[k], m_eager[k])
    assert float(d_eager.abs().max()) > 0.1 * cfg.learning_rate     # the update moved something
    # the first layers are padded to a multiple of 16 input features on CUDA: the brax parameter exchange must hide that
    import pickle
    assert agent._obs_pad == (-env.observation_size) % 16 and agent.policy[0].in_features == env.observation_size + agent._obs_pad
    norm, pol = agent.export_brax_params()
    assert pol["params"]["hidden_0"]["kernel"].shape == (env.observation_size, 32)
    other = PPO(env, PPOConfig(**{**cfg.__dict__, "seed": 9}))
    other.import_brax_params(pickle.loads(pickle.dumps((norm, pol))))
    a1, _, _ = agent.act(state.obs, deterministic=True)
    a2, _, _ = other.act(state.obs, deterministic=True)
    assert torch.allclose(a1, a2, atol=1e-5)
    assert float((d_graph - d_eager).abs().max()) < 0.05 * cfg.learning_rate, float((d_graph - d_eager).abs().max())


def _loss_case(emu_or_cuda_env, device, seed=0, T=5, B=37, tc=False):
    """Random minibatch incl. terminations / truncations, a few saturated ratios and a large pre-softplus scale."""
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    cfg = PPOConfig(**dict(TINY, num_envs=emu_or_cuda_env.num_envs, tf32=False, cuda_graph=False, tc_learner=tc))
    agent = PPO(emu_or_cuda_env, cfg)
    g = torch.Generator().manual_seed(seed)
    A, obs = emu_or_cuda_env.action_size, emu_or_cuda_env.observation_size
    r = lambda *s: torch.randn(*s, generator=g)
    obs += agent._obs_pad
    mb = dict(observation=r(T, B, obs), raw_action=1.5 * r(T, B, A), log_prob=-30 + 8 * r(T, B), reward=r(T, B),
              discount=(torch.rand(T, B, generator=g) > 0.2).float(), truncation=(torch.rand(T, B, generator=g) > 0.9).float(),
              next_observation_last=r(B, obs), entropy_noise=r(T, B, A))
    mb["truncation"] = mb["truncation"] * (1 - mb["discount"])          # truncation only where the episode ended
    mb = {k: v.to(device) for k, v in mb.items()}
    with torch.no_grad():
        agent.policy[-1].bias[A:A + 3] += 25.0                            # softplus threshold branch (> 20)
    agent._batch_is_normalized = True                                     # the random "observations" go straight into the MLPs
    return agent, mb


def _compare_fused_with_autograd(agent, mb, tol):
    import dataclasses
    outs = []
    for fused in (False, True):
        agent.cfg = dataclasses.replace(agent.cfg, fused_loss=fused)
        for p in agent.params:
            p.grad = None
        total, metrics = agent.loss(mb)
        total.backward()
        outs.append((float(total), {k: float(v) for k, v in metrics.items()},
                     torch.cat([p.grad.reshape(-1) for p in agent.params]).cpu()))
    (t0, m0, g0), (t1, m1, g1) = outs
    assert abs(t0 - t1) <= tol * max(1.0, abs(t0)), (t0, t1)
    for k in m0:
        assert abs(m0[k] - m1[k]) <= tol * max(1.0, abs(m0[k])), (k, m0[k], m1[k])
    assert float(g0.abs().max()) > 0
    assert float((g0 - g1).abs().max()) <= tol * float(g0.abs().max()), float((g0 - g1).abs().max())


def test_fused_loss_matches_autograd_on_emulator(emu_lib):
    agent, mb = _loss_case(tiny_env(emu_lib), "cpu")
    _compare_fused_with_autograd(agent, mb, 2e-4)


@pytest.mark.gpu
def test_fused_loss_matches_autograd_on_gpu():
    from brax_rodent_run_b200.env import Rodent
    env = Rodent(synthetic_track(), num_envs=2, device="cuda:0", model=load_asset("rodent_0"), iterations=1, ls_iterations=1, n_frames=1)
    agent, mb = _loss_case(env, "cuda:0", seed=1, T=10, B=512)
    _compare_fused_with_autograd(agent, mb, 2e-4)


class _BraxRunningStatisticsState:
    """Shape of brax.training.acme.running_statistics.RunningStatisticsState (a flax struct: attribute access, not a dict),
    as unpickled from a brax `model.save_params` file (brax_rodent_run_ppo.py:204-206, render_rollout.ipynb:117-118)."""

    def __init__(self, count, mean, summed_variance, std):
        self.count, self.mean, self.summed_variance, self.std = count, mean, summed_variance, std


def _brax_style_params(obs, act, hidden, seed):
    """A `(normalizer, policy)` tuple with brax's real structure: RunningStatisticsState-like object + flax param tree
    {"params": {"hidden_i": {"kernel": [in, out], "bias": [out]}}}, numpy leaves."""
    rng = np.random.default_rng(seed)
    sizes = (obs,) + tuple(hidden) + (2 * act,)
    tree = {"params": {f"hidden_{i}": {"kernel": rng.normal(0, 0.1, (sizes[i], sizes[i + 1])).astype(np.float32),
                                       "bias": rng.normal(0, 0.1, sizes[i + 1]).astype(np.float32)} for i in range(len(sizes) - 1)}}
    std = rng.uniform(0.5, 2.0, obs).astype(np.float32)
    norm = _BraxRunningStatisticsState(np.float32(1000.0), rng.normal(0, 1, obs).astype(np.float32), (std ** 2) * 1000.0, std)
    return norm, tree


def test_import_brax_structured_pickle_fixture(emu_lib, tmp_path):
    """A brax-structured checkpoint (object normaliser + flax tree, through pickle) drives the policy exactly as its own
    numpy forward pass: swish MLP on (obs - mean) / std, action = tanh(loc)."""
    import pickle
    from brax_rodent_run_b200 import ppo
    cfg = ppo.PPOConfig(**TINY)
    env = tiny_env(emu_lib).wrap_for_training(cfg.episode_length)
    params = _brax_style_params(env.observation_size, env.action_size, cfg.policy_hidden, 3)
    path = tmp_path / "brax_params.pkl"
    path.write_bytes(pickle.dumps(params))
    loaded = pickle.loads(path.read_bytes())
    agent = ppo.PPO(env, cfg)
    agent.import_brax_params(loaded)
    state = env.reset(1)
    obs = state.obs.numpy().astype(np.float64)
    h = (obs - params[0].mean) / params[0].std
    layers = params[1]["params"]
    for i in range(len(layers)):
        h = h @ layers[f"hidden_{i}"]["kernel"] + layers[f"hidden_{i}"]["bias"]
        if i + 1 < len(layers):
            h = h / (1 + np.exp(-h))
    want = np.tanh(h[:, :env.action_size])
    got, _, _ = agent.act(state.obs, deterministic=True)
    assert np.abs(got.numpy() - want).max() < 1e-5
    # the stand-alone make_policy(params, deterministic=True) of the training callback gives the same action
    pol = ppo.policy_from_brax_params(loaded, "cpu", deterministic=True)
    a2, extras = pol(state.obs, None)
    assert np.abs(a2.numpy() - want).max() < 1e-5 and set(extras) == {"log_prob", "raw_action"}


def test_train_callback_gets_make_policy_like_brax(emu_lib):
    """policy_params_fn(num_steps, make_policy, params) as the reference uses it (brax_rodent_run_ppo.py:135-151):
    make_policy(params, deterministic=True) -> inference_fn(obs, rng) -> (ctrl, extras); brax's evaluation cadence."""
    from brax_rodent_run_b200 import ppo
    cfg = ppo.PPOConfig(**dict(TINY, num_timesteps=48, num_evals=3))
    env = tiny_env(emu_lib)
    eval_env = tiny_env(emu_lib).wrap_for_training(cfg.episode_length)
    calls, progress = [], []

    def policy_params_fn(num_steps, make_policy, params):
        inference_fn = make_policy(params, deterministic=True)
        st = eval_env.reset(0)
        ctrl, _ = inference_fn(st.obs, None)
        st = eval_env.step(st, ctrl)
        calls.append((num_steps, tuple(ctrl.shape), bool(torch.isfinite(st.obs).all())))

    make_inference_fn, agent, metrics = ppo.train(env, cfg, progress_fn=lambda n, m: progress.append(n),
                                                  policy_params_fn=policy_params_fn, eval_env=eval_env)
    per_train = cfg.batch_size * cfg.num_minibatches * cfg.unroll_length
    # num_evals = 3 -> the untrained evaluation + 2 epochs of ceil(48 / (2 * 8)) = 3 training steps
    assert progress == [0, 3 * per_train, 6 * per_train]
    assert [c[0] for c in calls] == [3 * per_train, 6 * per_train] and all(c[1] == (2, env.action_size) and c[2] for c in calls)
    live = make_inference_fn()  # no params: the live agent
    a, extras = live(eval_env.reset(0).obs)
    assert a.shape == (2, env.action_size)


@pytest.mark.parametrize("tc", [False, True], ids=["autograd", "tc_learner"])
def test_resume_is_bitwise(emu_lib, tc):
    """Checkpoint / resume: state_dict -> load_state_dict into a fresh agent (same generator state, same env state) continues
    bit-for-bit: the next training step gives identical parameters and losses.  With the tensor-core learner the parameters are
    views of a flat buffer and the optimizer state is FlatAdam's."""
    import copy
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    cfg = PPOConfig(**dict(TINY, normalize_observations=False, tc_learner=tc))
    env = tiny_env(emu_lib).wrap_for_training(cfg.episode_length)
    a = PPO(env, cfg)
    state = env.reset(0)
    state, _ = a.training_step(state)
    ckpt = copy.deepcopy(a.state_dict())
    gen_state = a.gen.get_state()
    b = PPO(env, cfg)
    b.load_state_dict(ckpt)
    b.gen.set_state(gen_state)
    s1, m1 = a.training_step(state)
    s2, m2 = b.training_step(state)
    for p, q in zip(a.params, b.params):
        assert torch.equal(p, q)
    assert torch.equal(s1.obs, s2.obs) and all(float(m1[k]) == float(m2[k]) for k in m1)


@pytest.mark.gpu
def test_rollout_graph_matches_eager_unroll():
    """The captured unroll (policy MLP + tanh-normal sample + rr_step_kernel, unroll_length times, one CUDA graph) against the
    eager unroll with the same exploration noise: identical transitions and final state, over two consecutive replays."""
    from brax_rodent_run_b200.env import Rodent
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = PPOConfig(num_envs=64, batch_size=64, num_minibatches=2, unroll_length=5, episode_length=7)
    mk = lambda: Rodent(synthetic_track(), num_envs=64, device="cuda:0", model=load_asset("rodent_0"), iterations=2, ls_iterations=2,
                        n_frames=2, terminate_when_unhealthy=False, kinematics_outputs=False).wrap_for_training(cfg.episode_length)
    env_g, env_e = mk(), mk()
    a_g, a_e = PPO(env_g, cfg), PPO(env_e, PPOConfig(**dict(dataclasses_asdict(cfg), rollout_graph=False)))
    a_e.policy.load_state_dict(a_g.policy.state_dict())
    s_g, s_e = env_g.reset(3), env_e.reset(3)
    for rep in range(3):  # episode_length 7 < 15 steps: the fused auto-reset path is replayed too
        s_g, d_g = a_g.unroll(s_g)
        s_e, d_e = a_e._unroll_eager(s_e, a_g._rg_eps)  # same noise as the replay just used
        for k in d_g:
            assert torch.equal(d_g[k], d_e[k]), (rep, k)
        assert torch.equal(s_g.obs, s_e.obs) and torch.equal(s_g.pipeline_state.qpos, s_e.pipeline_state.qpos)
        assert torch.equal(s_g.info["steps"], s_e.info["steps"])


def dataclasses_asdict(cfg):
    import dataclasses
    return dataclasses.asdict(cfg)


def _compare_tc_with_autograd(agent, mb, tol):
    """loss_and_grads through the hand-written forward / backward (tc_learner.py on rr_tc_launch) against autograd."""
    outs = []
    flat_views = [p.grad for p in agent.params]      # the agent was built with tc_learner=True: views of the flat gradient buffer
    for use_tc in (False, True):
        agent._use_tc = use_tc
        for p, g in zip(agent.params, flat_views):
            p.grad = g if use_tc else None
        metrics = agent.loss_and_grads(mb)
        outs.append(({k: float(v) for k, v in metrics.items()}, torch.cat([p.grad.reshape(-1) for p in agent.params]).cpu().clone(),
                     [p.grad.detach().cpu().clone() for p in agent.params]))
    (m0, g0, l0), (m1, g1, l1) = outs
    for k in m0:
        assert abs(m0[k] - m1[k]) <= tol * max(1.0, abs(m0[k])), (k, m0[k], m1[k])
    assert float(g0.abs().max()) > 0
    for i, (a, b) in enumerate(zip(l0, l1)):     # per parameter tensor, relative to its own scale
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= tol * max(float(a.abs().max()), 1e-3 * float(g0.abs().max())), (i, a.shape)


def test_tc_learner_matches_autograd_on_emulator(emu_lib):
    agent, mb = _loss_case(tiny_env(emu_lib), "cpu", seed=3, T=3, B=5, tc=True)
    _compare_tc_with_autograd(agent, mb, 2e-4)
    assert agent._tc.launches_per_update == 3 + 2 + 1    # TINY: 3 forward layers, 2 dgrad steps, one wgrad launch
    assert agent._tc.splits == 1


def test_tc_learner_split_wgrad_on_emulator(emu_lib):
    """Enough rows for the weight gradients to be split four ways over the rows (partials + fixed-order sum)."""
    agent, mb = _loss_case(tiny_env(emu_lib), "cpu", seed=4, T=4, B=256, tc=True)
    _compare_tc_with_autograd(agent, mb, 5e-4)
    assert agent._tc.splits == 4


def test_training_step_with_tc_learner_on_emulator(emu_lib):
    """The whole training step (eager path on the CPU) with the tensor-core learner's launch lists on the emulator backend."""
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    outs = []
    for tc in (False, True):
        cfg = PPOConfig(**dict(TINY, tc_learner=tc))
        env = tiny_env(emu_lib).wrap_for_training(cfg.episode_length)
        agent = PPO(env, cfg)
        state = env.reset(0)
        state, metrics = agent.training_step(state)
        outs.append(torch.cat([p.detach().reshape(-1) for p in agent.params]))
    assert float((outs[0] - outs[1]).abs().max()) < 1e-4


@pytest.mark.gpu
def test_tc_learner_matches_autograd_on_gpu():
    """README minibatch shape (10 x 512 rows, 1264 features, 256-wide value net) through the tcgen05 kernel against fp32 autograd:
    TF32 products (10-bit mantissa, the weights truncated by the tensor core) through five layers forward and back, on a case
    built to have saturated ratios and gradients of 1e8 -- 3 % of each tensor's largest gradient."""
    from brax_rodent_run_b200.env import Rodent
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    env = Rodent(synthetic_track(), num_envs=2, device="cuda:0", model=load_asset("rodent_0"), iterations=1, ls_iterations=1, n_frames=1)
    cfg = PPOConfig(num_envs=2, batch_size=2, num_minibatches=2, unroll_length=2, cuda_graph=False, tf32=False, tc_learner=True)
    agent = PPO(env, cfg)
    _, mb = _loss_case(env, "cuda:0", seed=2, T=10, B=512)
    agent._batch_is_normalized = True
    _compare_tc_with_autograd(agent, mb, 3e-2)


@pytest.mark.gpu
def test_graphed_tc_update_matches_eager_autograd_update():
    """The captured update with the tensor-core learner applies (to TF32 accuracy) the same parameter update as the eager autograd
    update on the same minibatch, noise, parameters and Adam state."""
    from brax_rodent_run_b200.env import Rodent
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    cfg = PPOConfig(num_envs=64, batch_size=16, num_minibatches=8, unroll_length=4, num_updates_per_batch=2, episode_length=50,
                    num_timesteps=1, policy_hidden=(32, 32), value_hidden=(64, 64), tf32=False, cuda_graph=True, tc_learner=True)
    env = Rodent(synthetic_track(), num_envs=64, device="cuda:0", model=load_asset("rodent_0"), iterations=4, ls_iterations=4,
                 terminate_when_unhealthy=False).wrap_for_training(cfg.episode_length)
    agent = PPO(env, cfg)
    state = env.reset(0)
    state, metrics = agent.training_step(state)
    assert agent._graph is not None and agent._tc is not None
    assert all(math.isfinite(float(v)) for v in metrics.values())
    chunks = []
    for _ in range(cfg.batch_size * cfg.num_minibatches // cfg.num_envs):
        state, d = agent.unroll(state)
        chunks.append(d)
    data = {k: torch.cat([c[k] for c in chunks], dim=1 if k != "next_observation_last" else 0) for k in chunks[0]}
    data["observation"] = agent._norm(data["observation"])
    data["next_observation_last"] = agent._norm(data["next_observation_last"])
    agent._batch_is_normalized = True
    idx = torch.arange(3, 3 + cfg.batch_size, device="cuda:0")
    opt_tensors = [t for st in agent.opt.state.values() for t in st.values() if torch.is_tensor(t)]
    saved = [t.clone() for t in agent.params + opt_tensors]
    gen_state = agent.gen.get_state()
    m_graph = {k: float(v) for k, v in agent._update_graphed(data, idx).items()}
    d_graph = torch.cat([(p.detach() - v).reshape(-1) for p, v in zip(agent.params, saved)]).cpu()
    with torch.no_grad():
        for t, v in zip(agent.params + opt_tensors, saved):
            t.copy_(v)
    agent.gen.set_state(gen_state)
    st = agent._static
    st["entropy_noise"].normal_(generator=agent.gen)
    agent._use_tc = False
    tc_grads = [p.grad for p in agent.params]
    for p in agent.params:
        p.grad = None
    total, m = agent.loss(st)
    total.backward()
    for p, g in zip(agent.params, tc_grads):     # the optimizer steps over the flat gradient buffer
        g.copy_(p.grad)
        p.grad = g
    agent.opt.step()
    agent._use_tc = True
    d_eager = torch.cat([(p.detach() - v).reshape(-1) for p, v in zip(agent.params, saved)]).cpu()
    for k in m:
        assert abs(m_graph[k] - float(m[k])) <= 1e-2 * max(1.0, abs(float(m[k]))), (k, m_graph[k], float(m[k]))
    assert float(d_eager.abs().max()) > 0.1 * cfg.learning_rate
    # Adam's normalised step amplifies relative gradient error where the gradient is tiny: compare in the mean
    assert float((d_graph - d_eager).abs().mean()) < 0.05 * float(d_eager.abs().mean())


@pytest.mark.parametrize("backend", [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def test_fused_act_matches_torch_policy(backend, emu_lib):
    """rr_policy_act (normalise -> 1263->32x4->60 swish MLP -> tanh-normal sample + log-prob in one kernel) against the torch
    modules on the same parameters, normaliser and noise; stochastic and deterministic; then a whole unroll with it."""
    from brax_rodent_run_b200.env import Rodent
    from brax_rodent_run_b200.ppo import PPO, PPOConfig
    dev = "cpu" if backend == "emu" else "cuda:0"
    if backend == "cuda" and not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = 5 if backend == "emu" else 300
    kw = dict(_lib_path=emu_lib) if backend == "emu" else {}
    env = Rodent(synthetic_track(), num_envs=n, device=dev, model=load_asset("rodent_0"), iterations=1, ls_iterations=1, n_frames=1, **kw)
    cfg = PPOConfig(num_envs=n, batch_size=n, num_minibatches=1, unroll_length=2, value_hidden=(16,), fused_act=True, tf32=False,
                    cuda_graph=False, rollout_graph=False)
    agent = PPO(env.wrap_for_training(20), cfg)
    assert agent._use_fused_act
    g = torch.Generator().manual_seed(0)
    agent.normalizer.mean.copy_(torch.randn(env.observation_size, generator=g))
    agent.normalizer.std.copy_(torch.rand(env.observation_size, generator=g) + 0.5)
    obs = (2 * torch.randn(n, env.observation_size, generator=g)).to(dev)
    eps = torch.randn(n, env.action_size, generator=g).to(dev)
    if backend == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = False
    for deterministic in (False, True):
        agent._use_fused_act = True
        fused = agent.act(obs, deterministic, eps)
        agent._use_fused_act = False
        ref = agent.act(obs, deterministic, eps)
        for name, x, y, tol in zip(("action", "raw", "log_prob"), fused, ref, (2e-5, 1e-4, 2e-3)):
            assert float((x - y).abs().max()) <= tol, (deterministic, name, float((x - y).abs().max()))
    agent._use_fused_act = True
    state = env.reset(0) if backend == "cuda" else agent.env.reset(0)
    state, data = agent.unroll(state)
    assert torch.isfinite(data["log_prob"]).all() and data["raw_action"].shape == (2, n, env.action_size)
    with pytest.raises(ValueError):
        PPO(agent.env, PPOConfig(num_envs=n, batch_size=n, num_minibatches=1, policy_hidden=(8, 8), fused_act=True))
