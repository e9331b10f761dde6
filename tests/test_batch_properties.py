"""Size-independent properties of the batched step at BASELINE.json's full size (4096 envs on the GPU; a handful on the
emulator): environments are independent, so results must be bit-identical under re-execution, under any permutation of the
batch, under the explicit env -> warp-slot order of the load balancer, and across batch sizes (ragged last CTA, B = 1).
These pin the launch geometry (persistent CTAs, contiguous split, padding passes, rendezvous) -- bugs there do not show in
small-batch parity against the oracle."""
import numpy as np
import pytest
import torch

from conftest import backend_params, load_asset, synthetic_track

FIELDS = ("qpos", "qvel", "act", "qacc_warmstart")


def rollout(env, sf, nq_, nv_, actions):
    st = env.reset_from(torch.tensor(sf), torch.tensor(nq_), torch.tensor(nv_))
    for a in actions:
        st = env.step(st, torch.tensor(a))
    out = {k: getattr(st.pipeline_state, k).cpu().numpy().copy() for k in FIELDS}
    out.update(obs=st.obs.cpu().numpy().copy(), reward=st.reward.cpu().numpy().copy(), done=st.done.cpu().numpy().copy(),
               cur_frame=st.info["cur_frame"].cpu().numpy().copy())
    return out


def inputs(m, B, T, seed):
    rng = np.random.default_rng(seed)
    sf = rng.integers(0, 100, B)
    nq_, nv_ = rng.uniform(-.01, .01, (B, m.nq)), rng.uniform(-.01, .01, (B, m.nv))
    nq_[:, 2] += rng.uniform(-0.02, 0.03, B)  # some start in contact, some airborne: different solver work per warp
    acts = rng.uniform(-1, 1, (T, B, m.nu)).astype(np.float32)
    return sf, nq_, nv_, acts


@pytest.mark.parametrize("backend", backend_params())
def test_rerun_permutation_and_batch_size_invariance(backend, make_env):
    m, track = load_asset("rodent_0"), synthetic_track()
    B, T = (5, 2) if backend == "emu" else (4096, 3)
    kw = dict(model=m, iterations=4, ls_iterations=4, n_frames=2 if backend == "emu" else 10, terminate_when_unhealthy=True)
    sf, nq_, nv_, acts = inputs(m, B, T, 5)
    env = make_env(backend, track, num_envs=B, **kw).wrap_for_training(episode_length=1000)
    ref = rollout(env, sf, nq_, nv_, acts)
    assert np.isfinite(ref["obs"]).all()
    np.testing.assert_allclose(np.linalg.norm(ref["qpos"][:, 3:7], axis=1), 1.0, atol=2e-6)
    # 1. re-execution
    again = rollout(env, sf, nq_, nv_, acts)
    for k in ref:
        assert np.array_equal(ref[k], again[k]), k
    # 2. permutation of the batch
    perm = np.random.default_rng(9).permutation(B)
    p = rollout(env, sf[perm], nq_[perm], nv_[perm], acts[:, perm])
    for k in ref:
        assert np.array_equal(ref[k][perm], p[k]), k
    # 3. a smaller, ragged batch (and a single environment) reproduces the same environments
    for Bs in ((3, 1) if backend == "emu" else (1481, 37, 1)):
        small = make_env(backend, track, num_envs=Bs, **kw).wrap_for_training(episode_length=1000)
        s = rollout(small, sf[:Bs], nq_[:Bs], nv_[:Bs], acts[:, :Bs])
        for k in ref:
            assert np.array_equal(ref[k][:Bs], s[k]), (Bs, k)


@pytest.mark.gpu
def test_balanced_order_is_bitwise_neutral():
    """Rodent(balance=True) regroups environments over CTAs by last step's solver work; results must not depend on it."""
    from brax_rodent_run_b200.env import Rodent
    m, track = load_asset("rodent_0"), synthetic_track()
    B, T = 4096, 3
    sf, nq_, nv_, acts = inputs(m, B, T, 6)
    outs = []
    for balance in (False, True):
        env = Rodent(track, num_envs=B, device="cuda:0", model=m, iterations=4, ls_iterations=4, balance=balance,
                     terminate_when_unhealthy=True).wrap_for_training(episode_length=1000)
        assert env._balance == balance
        outs.append(rollout(env, sf, nq_, nv_, acts))
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_emulator_lane_order_independence(emu_lib, make_env, monkeypatch):
    """Race check on the host (compute-sanitizer is closed on the GPU pool, profiles/r02_sanitizer_closed.txt): the fiber
    emulator runs the 32 lanes of a warp round-robin between warp primitives; running them in reverse and in odd-first order
    must give bit-identical results.  A shared-memory hazard between lanes that no __syncwarp / shuffle / ballot separates
    shows up as a difference (the emulator also poisons shared memory with NaN before every environment)."""
    import os
    m = load_asset("rodent_0")
    rng = np.random.default_rng(3)
    B = 2
    sf, nq_, nv_ = rng.integers(0, 100, B), rng.uniform(-.01, .01, (B, m.nq)), rng.uniform(-.01, .01, (B, m.nv))
    act = rng.uniform(-1, 1, (2, B, m.nu)).astype(np.float32)
    outs = []
    for order in (None, "reverse", "odd-first"):
        if order is None:
            monkeypatch.delenv("RR_EMU_LANE_ORDER", raising=False)
        else:
            monkeypatch.setenv("RR_EMU_LANE_ORDER", order)
        env = make_env("emu", synthetic_track(), num_envs=B, model=m, iterations=4, ls_iterations=4, n_frames=3).wrap_for_training(1000)
        st = env.reset_from(torch.tensor(sf), torch.tensor(nq_), torch.tensor(nv_))
        for t in range(2):
            st = env.step(st, torch.tensor(act[t]))
        outs.append((st.obs.numpy().copy(), st.pipeline_state.qpos.numpy().copy(), st.pipeline_state.qacc_warmstart.numpy().copy()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)


def test_reuse_buffers_matches_fresh_allocation(make_env):
    """Rodent(reuse_buffers=True) cycles three preallocated output sets; a `s = env.step(s, a)` loop gives bit-identical states."""
    m = load_asset("rodent_0")
    rng = np.random.default_rng(5)
    acts = rng.uniform(-1, 1, (5, 2, m.nu)).astype(np.float32)
    outs = []
    for reuse in (False, True):
        env = make_env("emu", synthetic_track(), num_envs=2, model=m, iterations=2, ls_iterations=2, n_frames=1,
                       reuse_buffers=reuse).wrap_for_training(3)
        s = env.reset(4)
        prev = None
        for t in range(5):
            prev, s = s, env.step(s, torch.tensor(acts[t]))
            assert prev.obs.data_ptr() != s.obs.data_ptr()  # the previous state is still intact
        outs.append((s.obs.numpy().copy(), s.pipeline_state.qpos.numpy().copy(), s.info["steps"].numpy().copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
