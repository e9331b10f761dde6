"""MJCF loader (brax_rodent_run_b200/mjcf.py): compiled assets, re-parse of the reference's XML files when they are
present (this container only -- /root/reference does not exist on the GPU box), defaults / error behaviour."""
import os

import numpy as np
import pytest

from conftest import ROOT, load_asset

REF_MODELS = "/root/reference/models"


@pytest.mark.parametrize("name", ["rodent_0", "rodent_new", "rodent_optimized", "rodent_pair"])
def test_assets_match_reparse(name):
    from brax_rodent_run_b200 import mjcf
    path = os.path.join(REF_MODELS, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference models not available here")
    a, b = load_asset(name), mjcf.load_xml(path)
    assert (a.nq, a.nv, a.nu, a.nbody, a.ngeom, a.ncon, a.nefc) == (b.nq, b.nv, b.nu, b.nbody, b.ngeom, b.ncon, b.nefc)
    for k in a.arrays:
        assert np.allclose(a.arrays[k], b.arrays[k], rtol=1e-12, atol=1e-14), k
    assert abs(a.meaninertia - b.meaninertia) < 1e-14


def test_freejoint_ignores_defaults(rodent0):
    """<freejoint> takes no defaults: armature 0 / damping 0 on the root dofs."""
    assert np.all(rodent0.dof_armature[:6] == 0) and np.all(rodent0.dof_damping[:6] == 0)
    assert np.all(rodent0.dof_armature[6:] > 0) and np.all(rodent0.dof_damping[6:] > 0)


def test_paw_priority_mixing(rodent0):
    """Paw capsules have priority 1 over the floor's 0: the contact takes the paw's friction / solref (rodent_0.xml:31-45)."""
    m = rodent0
    mu = m.pair_friction
    assert set(np.round(mu, 6)) <= {1.5, 0.7}
    assert (mu == 1.5).sum() >= 8


def test_actuators(rodent0):
    m = rodent0
    assert np.all(m.actuator_dyntype == 2) and np.all(m.actuator_biastype == 1)  # filter dynamics, affine bias
    assert np.allclose(m.actuator_dynprm, 0.04)
    assert np.all(m.actuator_ctrllimited == 1) and np.all(m.actuator_forcelimited == 0)  # forcerange is ignored
    assert np.allclose(m.actuator_ctrlrange, np.tile([-1.0, 1.0], (m.nu, 1)))


def test_tree_layout(rodent0):
    m = rodent0
    depth = m.M_rownnz - 1
    assert depth.max() == 35 and m.M_rownnz.sum() == m.nM
    # 6 root dofs chained, then every chain hangs off dof 5 (neck, arms) or dof 11 (legs, tail)
    assert m.dof_parentid[:12].tolist() == [-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
    assert sorted(set(m.dof_parentid[[12, 18, 24]])) == [11] and sorted(set(m.dof_parentid[[48, 57, 65]])) == [5]


def test_bad_xml(tmp_path):
    from brax_rodent_run_b200 import mjcf
    p = tmp_path / "bad.xml"
    p.write_text("<mujoco><worldbody><body><joint type='ball'/><geom size='0.1'/></body></worldbody></mujoco>")
    with pytest.raises(NotImplementedError):
        mjcf.load_xml(str(p))
    p.write_text("<notmujoco/>")
    with pytest.raises(ValueError):
        mjcf.load_xml(str(p))
    p.write_text("<mujoco><worldbody><body><joint/><geom size='0.1' class='nope'/></body></worldbody></mujoco>")
    with pytest.raises(ValueError):
        mjcf.load_xml(str(p))


def test_minimal_model_roundtrip(tmp_path, emu_lib, oracle_mod):
    """A 2-link pendulum parsed from XML runs through the C ABI (emulator) and matches the oracle: the loader + kernel are
    not specialised to the rodent topology."""
    import torch
    from brax_rodent_run_b200 import mjcf, model_blob
    from brax_rodent_run_b200.env import Rodent
    p = tmp_path / "pend.xml"
    p.write_text("""<mujoco><compiler angle="radian"/><worldbody>
      <geom name="floor" type="plane" size="1 1 .1" conaffinity="1" contype="0"/>
      <body name="a" pos="0 0 .3"><freejoint/><geom type="sphere" size=".05" contype="1" conaffinity="0"/>
        <body name="b" pos=".1 0 0"><joint name="j" axis="0 1 0" range="-1 1" damping=".01" armature=".001"/>
          <geom type="capsule" size=".02 .05" contype="1" conaffinity="0"/></body></body></worldbody>
      <actuator><general joint="j" dyntype="filter" dynprm=".04" gainprm="1" biastype="affine" biasprm="0 -1"
        ctrllimited="true" ctrlrange="-1 1"/></actuator></mujoco>""")
    m = mjcf.load_xml(str(p))
    assert (m.nq, m.nv, m.nu, m.nbody, m.ncon) == (8, 7, 1, 3, 3)
    env = Rodent(np.zeros((4, 3), np.float32), device="cpu", xml_path=str(p), iterations=4, ls_iterations=4, n_frames=1,
                 _lib_path=emu_lib)
    qpos = torch.tensor(m.qpos0[None], dtype=torch.float32)
    qpos[0, 2] = 0.04
    qpos[0, 7] = 1.2  # beyond the hinge limit
    out = env.debug_forward(qpos, torch.full((1, m.nv), 0.1))
    o = oracle_mod.Oracle(model_blob.pack(m), "f64")
    o.set_options(0, 4, 4)
    o.init(qpos[0].numpy().astype(np.float64), np.full(m.nv, 0.1))
    for k in ("xpos", "cinert", "contact_dist", "efc_J", "qacc"):
        a, b = out[k][0].numpy(), o.get(k)
        assert np.abs(a - b).max() / (np.abs(b).max() + 1e-30) < 1e-4, k
