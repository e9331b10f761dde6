"""Reference-clip pipeline (preprocessing/mjx_preprocess.py restated in brax_rodent_run_b200/preprocess.py): the batched
forward kinematics against the CPU oracle, the finite-difference velocities against closed forms, the 0.9 rescale against the
scaling law of a kinematic tree, and the storage round trips."""
import os
import pickle

import numpy as np
import pytest

from conftest import ROOT, backend_params, has_cuda, load_asset


def _dev(backend, emu_lib):
    if backend == "emu":
        return dict(device="cpu", _lib_path=emu_lib)
    if not has_cuda():
        pytest.skip("no CUDA device")
    return dict(device="cuda:0")


def _mocap(m, T, seed=0):
    rng = np.random.default_rng(seed)
    q = np.tile(m.qpos0, (T, 1)) + np.cumsum(rng.uniform(-.02, .02, (T, m.nq)), 0)
    q[:, 3:7] = np.array([1.0, 0, 0, 0]) + np.cumsum(rng.uniform(-.05, .05, (T, 4)), 0)  # un-normalised, as STAC output can be
    return q.astype(np.float32)


@pytest.mark.parametrize("backend", backend_params())
def test_extract_features_matches_oracle_kinematics(backend, emu_lib, oracle_mod):
    from brax_rodent_run_b200 import model_blob, preprocess
    m = load_asset("rodent_0")
    T = 4 if backend == "emu" else 64
    q = _mocap(m, T)
    clip = preprocess.process_clip(q, m, **_dev(backend, emu_lib))
    assert clip.position.shape == (T, 3) and clip.quaternion.shape == (T, 4) and clip.joints.shape == (T, m.nq - 7)
    assert clip.body_positions.shape == (T, m.nbody, 3) and clip.body_quaternions.shape == (T, m.nbody, 4)
    assert clip.velocity.shape == (T, 3) and clip.angular_velocity.shape == (T, 3) and clip.joints_velocity.shape == (T, m.nv - 6)
    np.testing.assert_allclose(np.linalg.norm(clip.quaternion, axis=1), 1.0, atol=1e-6)   # written back normalised
    np.testing.assert_array_equal(clip.position, q[:, :3])
    np.testing.assert_array_equal(clip.joints, q[:, 7:])
    blob = model_blob.pack(m)
    for t in range(min(T, 4)):
        o = oracle_mod.Oracle(blob, "f64")
        o.init(q[t].astype(np.float64), np.zeros(m.nv))
        np.testing.assert_allclose(clip.body_positions[t].ravel(), o.get("xpos"), atol=2e-6)
        np.testing.assert_allclose(clip.body_quaternions[t].ravel(), o.get("xquat"), atol=2e-6)


def test_velocity_from_kinematics_closed_form():
    from brax_rodent_run_b200 import preprocess as pp
    dt, T = 0.02, 6
    axis = np.array([1.0, 2.0, -0.5]); axis /= np.linalg.norm(axis)
    w = 3.0                                                  # rad/s about `axis`, in the body frame
    t = np.arange(T) * dt
    quat = np.concatenate([np.cos(w * t / 2)[:, None], np.sin(w * t / 2)[:, None] * axis], 1)
    q = np.zeros((T, 7 + 3))
    q[:, :3] = np.outer(t, [0.5, -1.0, 2.0])
    q[:, 3:7] = quat * 1.7                                   # not unit: the difference is normalised before the axis-angle
    q[:, 7:] = np.outer(t, [1.0, -2.0, 40.0 * 50])
    v = pp.compute_velocity_from_kinematics(q, dt)
    assert v.shape == (T - 1, 9)
    np.testing.assert_allclose(v[:, :3], np.tile([0.5, -1.0, 2.0], (T - 1, 1)), atol=1e-12)
    np.testing.assert_allclose(v[:, 3:6], np.tile(w * axis, (T - 1, 1)), atol=1e-9)
    np.testing.assert_allclose(v[:, 6:], np.tile([1.0, -2.0, 2000.0], (T - 1, 1)), atol=1e-9)
    # identical consecutive quaternions: zero rotation, no NaN from the 0 / 0 axis (transformations.py: angle < 1e-10 -> zeros)
    q2 = np.zeros((3, 8)); q2[:, 3] = 1.0
    assert np.array_equal(pp.compute_velocity_from_kinematics(q2, dt), np.zeros((2, 7)))
    # quat_to_axisangle wraps the angle into [-pi, pi): a rotation by 1.5 pi comes back as -0.5 pi about the same axis
    big = np.array([np.cos(0.75 * np.pi), np.sin(0.75 * np.pi), 0.0, 0.0])
    np.testing.assert_allclose(pp.quat_to_axisangle(big), [-0.5 * np.pi, 0.0, 0.0], atol=1e-12)


def test_process_clip_pads_and_clips(emu_lib):
    from brax_rodent_run_b200 import preprocess as pp
    m = load_asset("rodent_0")
    q = np.tile(m.qpos0, (3, 1)).astype(np.float32)
    q[1, 7] += 1.0                                           # 1 rad in 20 ms = 50 rad/s -> clipped to max_qvel
    q[2, 7] += 1.0
    q[1, 0] += 1.0                                           # the root's linear velocity is not clipped
    clip = pp.process_clip(q, m, max_qvel=20.0, dt=0.02, device="cpu", _lib_path=emu_lib)
    assert clip.joints_velocity[0, 0] == 20.0 and clip.joints_velocity[1, 0] == 0.0
    assert abs(clip.velocity[0, 0] - 50.0) < 1e-3 and abs(clip.velocity[1, 0] + 50.0) < 1e-3
    assert not clip.velocity[2].any() and not clip.joints_velocity[2].any() and not clip.angular_velocity[2].any()  # padding


_TINY_XML = """<mujoco><compiler angle="radian"/>
<default><joint limited="true" range="-1 1" armature="0.001" damping="0.01"/><geom density="500" contype="1" conaffinity="0"/></default>
<worldbody><geom name="floor" type="plane" size="1 1 .1" pos="0 0 -0.01" contype="0" conaffinity="1"/>
 <body name="torso" pos="0 0 0.1"><freejoint/><geom type="capsule" size="0.02 0.05" pos="0.01 0 0" quat="0.7071 0 0.7071 0"/>
  <body name="a" pos="0.05 0.01 0"><joint type="hinge" axis="0 1 0" pos="0.01 0 0"/><geom type="sphere" size="0.01" pos="0.02 0 0"/>
   <body name="b" pos="0.03 0 -0.02" quat="0.9 0.1 0 0.2"><joint type="hinge" axis="1 0 0"/><geom type="capsule" size="0.005 0.02"/></body>
  </body>
 </body>
</worldbody></mujoco>"""


def _rescale_case(xml, emu_lib, tmp_path, exact=True):
    from brax_rodent_run_b200 import mjcf, preprocess as pp
    m1, m9 = mjcf.load_xml(xml), mjcf.load_xml(xml, rescale=(0.9, 0.9))
    np.testing.assert_allclose(m9.body_pos, 0.9 * m1.body_pos, rtol=1e-12)
    if exact:
        np.testing.assert_allclose(m9.jnt_pos, 0.9 * m1.jnt_pos, rtol=1e-12)
    rng = np.random.default_rng(3)
    q = np.tile(m1.qpos0, (6, 1)) + rng.uniform(-.2, .2, (6, m1.nq))
    q[:, :3] = 0.0
    q[:, 3:7] = [1.0, 0, 0, 0]
    stac = tmp_path / "stac.p"
    with open(stac, "wb") as f:
        pickle.dump({"qpos": q}, f)
    c9 = pp.process_clip_to_train(str(stac), xml, scale_factor=0.9, start_step=1, clip_length=4, device="cpu", _lib_path=emu_lib)
    c1 = pp.process_clip(q[1:5], m1, device="cpu", _lib_path=emu_lib)
    assert c9.body_positions.shape[0] == 4
    if exact:
        np.testing.assert_allclose(c9.body_positions, 0.9 * c1.body_positions, atol=2e-6)
    np.testing.assert_allclose(c9.body_quaternions, c1.body_quaternions, atol=2e-6)
    return m1, m9, c1, c9


def test_rescaled_model_scales_the_tree(emu_lib, tmp_path):
    """dm_control rescale_subtree(0.9, 0.9): with the root at the origin every body position scales by exactly 0.9 and no
    orientation changes; process_clip_to_train reads the STAC pickle window [start_step, start_step + clip_length)."""
    xml = tmp_path / "tiny.xml"
    xml.write_text(_TINY_XML)
    m1, m9, _, _ = _rescale_case(str(xml), emu_lib, tmp_path)
    # explicit sizes scale too (capsule radius and half-length), masses follow the volumes
    np.testing.assert_allclose(m9.geom_size[1:, 0], 0.9 * m1.geom_size[1:, 0], rtol=1e-12)
    np.testing.assert_allclose(m9.body_mass[1:], 0.9 ** 3 * m1.body_mass[1:], rtol=1e-9)


def test_rescaled_reference_rodent(emu_lib, tmp_path):
    xml = "/root/reference/models/rodent_0.xml"
    if not os.path.exists(xml):
        pytest.skip("reference models not available here")
    # joint anchors that come from <default> classes keep their size (rescale_subtree only touches explicit attributes), so
    # the tree is not an exact 0.9 copy: a few of the 204 anchor coordinates stay, and body positions agree to those 0.3 mm
    m1, m9, c1, c9 = _rescale_case(xml, emu_lib, tmp_path, exact=False)
    same = np.isclose(m9.jnt_pos, m1.jnt_pos) & (m1.jnt_pos != 0)
    assert 0 < same.sum() < 0.2 * m1.jnt_pos.size
    assert np.abs(c9.body_positions - 0.9 * c1.body_positions).max() < 2e-3


def test_clip_storage_round_trip(emu_lib, tmp_path):
    from brax_rodent_run_b200 import preprocess as pp
    from brax_rodent_run_b200.run_ppo import load_track
    m = load_asset("rodent_0")
    clip = pp.process_clip(_mocap(m, 3), m, device="cpu", _lib_path=emu_lib)
    path = str(tmp_path / "clip.p")
    pp.save_reference_clip(path, clip)
    back = pp.load_reference_clip(path)
    for k, v in vars(clip).items():
        np.testing.assert_array_equal(getattr(back, k), v)
    np.testing.assert_array_equal(load_track(path), clip.position)      # what the training entry point reads (track_pos)
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            pp.save_reference_clip_to_h5(str(tmp_path / "clip.h5"), "clip_0", clip)
    else:
        pp.save_reference_clip_to_h5(str(tmp_path / "clip.h5"), "clip_0", clip)
        h = pp.load_reference_clip_from_h5(str(tmp_path / "clip.h5"), "clip_0")
        np.testing.assert_array_equal(h.position[0], clip.position)
