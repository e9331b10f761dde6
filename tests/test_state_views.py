"""brax.mjx State views on pipeline_state: x, xd, contact -- what Rodent_Env_Brax.py:101 returns
through brax.mjx.pipeline.step and torchrl_explore.ipynb:609-618 reads -- checked against the oracle's mjx.Data."""
import numpy as np
import pytest
import torch

from conftest import backend_params, load_asset, synthetic_track
from test_parity_step import draws, oracle_env, rel


@pytest.mark.parametrize("backend", backend_params())
def test_x_xd_contact_views(backend, make_env, oracle_mod):
    m, track = load_asset("rodent_0"), synthetic_track()
    B = 2
    kw = dict(iterations=4, ls_iterations=4, n_frames=1)
    env = make_env(backend, track, num_envs=B, model=m, **kw)
    sf, nq_, nv_ = draws(m, B, 13)
    nv_ *= 50  # visible velocities
    st = env.reset_from(torch.tensor(sf), torch.tensor(nq_), torch.tensor(nv_))
    for e in range(B):
        oe = oracle_env(oracle_mod, m, track, "f64", **kw)
        q = m.qpos0.copy()
        q[:3] = track[sf[e]]
        oe.reset(sf[e], q + nq_[e], nv_[e])
        o = oe.o
        ps = st.pipeline_state
        xpos, xquat = o.get("xpos").reshape(-1, 3), o.get("xquat").reshape(-1, 4)
        assert ps.x.pos.shape == (B, m.nbody - 1, 3) and ps.x.rot.shape == (B, m.nbody - 1, 4)
        assert rel(ps.x.pos[e].cpu().numpy(), xpos[1:]) < 2e-5 and rel(ps.x.rot[e].cpu().numpy(), xquat[1:]) < 2e-5
        # xd: cvel moved from the subtree COM of the tree root to the link origin
        cvel = o.get("cvel").reshape(-1, 6)
        com = o.get("subtree_com").reshape(-1, 3)
        off = xpos[1:] - com[np.asarray(m.body_rootid)[1:]]
        vel = cvel[1:, 3:] - np.cross(off, cvel[1:, :3])
        assert rel(ps.xd.ang[e].cpu().numpy(), cvel[1:, :3]) < 2e-5 and rel(ps.xd.vel[e].cpu().numpy(), vel) < 5e-5
        # the free-joint root: xd of the torso is its world linear velocity = qvel[:3]
        assert np.abs(ps.xd.vel[e, 0].cpu().numpy() - nv_[e, :3]).max() < 1e-4 * max(1.0, np.abs(nv_[e, :3]).max())
        c = ps.contact
        assert c.dist.shape == (B, m.ncon) and c.pos.shape == (B, m.ncon, 3) and c.frame.shape == (B, m.ncon, 3, 3)
        assert rel(c.dist[e].cpu().numpy(), o.get("contact_dist")) < 2e-5
        assert rel(c.pos[e].cpu().numpy(), o.get("contact_pos")) < 2e-5
        assert rel(c.frame[e].cpu().numpy(), o.get("contact_frame")) < 2e-5
        # contact ordering / typing: capsule pairs first (two contacts each), floor is geom1 -> link_idx[0] = -1 (world)
        assert (c.link_idx[0].cpu().numpy() == -1).all() and (c.link_idx[1].cpu().numpy() >= 0).all()
        assert c.friction.shape == (m.ncon, 5) and c.solref.shape == (m.ncon, 2) and c.solimp.shape == (m.ncon, 5)
        assert c.geom1.shape == (m.ncon,) and float(c.elasticity.abs().max()) == 0.0
    # the views follow the step
    st2 = env.step(st, torch.zeros(B, m.nu))
    assert st2.pipeline_state.contact.dist.shape == (B, m.ncon) and torch.isfinite(st2.pipeline_state.xd.vel).all()


def test_views_need_kinematics_outputs(make_env):
    m = load_asset("rodent_0")
    env = make_env("emu", synthetic_track(), num_envs=1, model=m, iterations=2, ls_iterations=2, n_frames=1, kinematics_outputs=False)
    st = env.reset(0)
    for name in ("x", "xd", "contact", "xmat"):
        with pytest.raises(AttributeError):
            getattr(st.pipeline_state, name)
