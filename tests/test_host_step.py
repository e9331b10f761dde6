"""rr_env_step_host (the end-to-end entry point: host action in, host obs / reward / done out) against the device-buffer
step on the same state: bit-identical, through the zero-copy path (pinned obs buffer) and the copy path (pageable)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import backend_params, load_asset, synthetic_track


@pytest.mark.parametrize("backend", backend_params())
def test_host_step_matches_device_step(backend, make_env):
    from brax_rodent_run_b200 import _lib
    m, track = load_asset("rodent_0"), synthetic_track()
    B = 3 if backend == "emu" else 300
    env = make_env(backend, track, num_envs=B, model=m, iterations=4, ls_iterations=4, n_frames=2 if backend == "emu" else 10)
    env.wrap_for_training(episode_length=1000)
    L = env._L
    st = env.reset(3)
    rng = np.random.default_rng(0)
    act = torch.tensor(rng.uniform(-1, 1, (B, m.nu)).astype(np.float32))
    ref = env.step(st, act.to(env.device))                          # device path (fresh output buffers)
    for pinned in ((False, True) if backend == "cuda" else (False,)):
        buf, t = env._out_buffers()
        ps = st.pipeline_state
        for k, v in (("qpos", ps.qpos), ("qvel", ps.qvel), ("act", ps.act), ("qacc_warmstart", ps.qacc_warmstart), ("time", ps.time),
                     ("cur_frame", st.info["cur_frame"]), ("done", st.done), ("steps", st.info["steps"])):
            t[k].copy_(v)
        f = st.info["first_pipeline_state"]
        buf.first_qpos, buf.first_qvel, buf.first_act = f.qpos.data_ptr(), f.qvel.data_ptr(), f.act.data_ptr()
        buf.first_qacc_warmstart, buf.first_time, buf.first_obs = f.qacc_warmstart.data_ptr(), f.time.data_ptr(), st.info["first_obs"].data_ptr()
        t["obs"].fill_(-7.0)
        h_act = act.clone().pin_memory() if pinned else act.clone()
        h_obs = torch.empty((B, env.observation_size), pin_memory=pinned)
        h_rew, h_done = torch.empty((B,), pin_memory=pinned), torch.empty((B,), pin_memory=pinned)
        p = lambda x: ctypes.c_void_p(x.data_ptr())
        _lib.check(L, L.rr_env_step_host(env._env, ctypes.byref(buf), p(h_act), env.n_frames, p(h_obs), p(h_rew), p(h_done), env._stream()))
        assert torch.equal(h_obs, ref.obs.cpu()), pinned
        assert torch.equal(h_rew, ref.reward.cpu()) and torch.equal(h_done, ref.done.cpu())
        assert torch.equal(t["qpos"].cpu(), ref.pipeline_state.qpos.cpu())
        if pinned:   # zero-copy: the device observation buffer is documented as untouched
            assert float(t["obs"].min()) == -7.0 and float(t["obs"].max()) == -7.0
        else:
            assert torch.equal(t["obs"].cpu(), ref.obs.cpu())
