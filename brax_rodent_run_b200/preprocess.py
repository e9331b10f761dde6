"""Reference-clip pipeline: mocap joint angles -> the `ReferenceClip` the run task tracks.

Mirrors preprocessing/mjx_preprocess.py of the reference (same names, arguments, field layout, clipping and padding):
  ReferenceClip                      (:23-41)   position / quaternion / joints / body_positions / body_quaternions +
                                                finite-difference velocity / angular_velocity / joints_velocity
  process_clip_to_train              (:44-91)   STAC pickle {"qpos": [T, nq]} + MJCF rescaled by 0.9 -> ReferenceClip
  process_clip                       (:94-135)
  extract_features / set_position    (:138-195) forward kinematics of every frame
  compute_velocity_from_kinematics   (:198-222) with preprocessing/transformations.py quat_diff / quat_to_axisangle
  save_/load_reference_clip_*                   pickle always; HDF5 (:225-283) when h5py is importable

The forward kinematics of all T frames run as ONE batch through the CUDA step kernel's forward mode (`rr_env_init`, one
warp per frame) -- the same kinematics code the physics step uses -- instead of the reference's `lax.scan` over frames.
There is no CPU fallback: the model handle lives on a CUDA device (tests drive the same code through the emulator).
"""
from __future__ import annotations

import dataclasses
import pickle
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import mjcf
from .env import Rodent

_TOL = 1e-10  # transformations.py:8


@dataclasses.dataclass
class ReferenceClip:
    """preprocessing/mjx_preprocess.py:23-41 (a flax struct there; `.replace` kept)."""
    position: Optional[np.ndarray] = None          # [T, 3]   qpos[:3]
    quaternion: Optional[np.ndarray] = None        # [T, 4]   qpos[3:7], normalised by the kinematics
    joints: Optional[np.ndarray] = None            # [T, nq-7]
    body_positions: Optional[np.ndarray] = None    # [T, nbody, 3]  xpos
    velocity: Optional[np.ndarray] = None          # [T, 3]
    joints_velocity: Optional[np.ndarray] = None   # [T, nv-6]
    angular_velocity: Optional[np.ndarray] = None  # [T, 3]
    body_quaternions: Optional[np.ndarray] = None  # [T, nbody, 4]  xquat

    def replace(self, **kw) -> "ReferenceClip":
        return dataclasses.replace(self, **kw)


# ---- preprocessing/transformations.py -------------------------------------------------------------------------------
def quat_mul(q1: np.ndarray, q2: np.ndarray) -> np.ndarray:
    """Hamilton product, any leading batch dimensions (transformations.py:30-52)."""
    w1, x1, y1, z1 = np.moveaxis(np.asarray(q1), -1, 0)
    w2, x2, y2, z2 = np.moveaxis(np.asarray(q2), -1, 0)
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)


def quat_conj(q: np.ndarray) -> np.ndarray:
    q = np.asarray(q)
    return np.concatenate([q[..., :1], -q[..., 1:]], -1)


def quat_diff(source: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Rotation from source to target (transformations.py:103-116)."""
    return quat_mul(quat_conj(source), target)


def quat_to_axisangle(quat: np.ndarray) -> np.ndarray:
    """Axis * angle of a unit quaternion, batched (transformations.py:119-139): angle wrapped to [-pi, pi), zero below 1e-10."""
    quat = np.asarray(quat)
    angle = 2.0 * np.arccos(np.clip(quat[..., 0], -1.0, 1.0))
    small = angle < _TOL
    qn = np.where(small, 1.0, np.sin(angle / 2.0))
    wrapped = (angle + np.pi) % (2.0 * np.pi) - np.pi
    out = quat[..., 1:4] / qn[..., None] * wrapped[..., None]
    return np.where(small[..., None], 0.0, out)


def compute_velocity_from_kinematics(qpos_trajectory: np.ndarray, dt: float) -> np.ndarray:
    """[T, nq] -> [T-1, nq-1] finite-difference velocities; free joint in the first 7 entries (mjx_preprocess.py:198-222)."""
    q = np.asarray(qpos_trajectory)
    lin = (q[1:, :3] - q[:-1, :3]) / dt
    d = quat_diff(q[:-1, 3:7], q[1:, 3:7])
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    gyro = quat_to_axisangle(d) / dt
    joints = (q[1:, 7:] - q[:-1, 7:]) / dt
    return np.concatenate([lin, gyro, joints], axis=1)


# ---- forward kinematics over the clip -------------------------------------------------------------------------------
def extract_features(model: mjcf.FlatModel, clip: ReferenceClip, mocap_qpos, device="cuda:0", _lib_path=None) -> ReferenceClip:
    """set_position + kinematics of every frame (mjx_preprocess.py:138-195), all frames in one kernel launch."""
    q = torch.as_tensor(np.asarray(mocap_qpos), dtype=torch.float32)
    T = q.shape[0]
    if q.ndim != 2 or q.shape[1] != model.nq:
        raise ValueError(f"mocap_qpos must be [T, {model.nq}], got {tuple(q.shape)}")
    fk = Rodent(np.zeros((1, 3), np.float32), num_envs=T, device=device, model=model, kinematics_outputs=True, _lib_path=_lib_path)
    st = fk.init_state(q, torch.zeros((T, model.nv)))
    ps = st.pipeline_state
    qpos = ps.qpos.cpu().numpy()  # root quaternion normalised by the kinematics, as mjx_data.qpos after mjx kinematics
    return clip.replace(position=qpos[:, :3], quaternion=qpos[:, 3:7], joints=qpos[:, 7:],
                        body_positions=ps.xpos.cpu().numpy(), body_quaternions=ps.xquat.cpu().numpy())


def process_clip(mocap_qpos, model: mjcf.FlatModel, max_qvel: float = 20.0, dt: float = 0.02, device="cuda:0",
                 _lib_path=None) -> ReferenceClip:
    """mjx_preprocess.py:94-135 (the reference passes mjx_model / mjx_data; here the flat model)."""
    mocap_qpos = np.asarray(mocap_qpos, np.float32)
    clip = extract_features(model, ReferenceClip(), mocap_qpos, device=device, _lib_path=_lib_path)
    padded = np.concatenate([mocap_qpos, mocap_qpos[-1:]], axis=0)  # velocity corner case: last frame repeated -> zero velocity
    qvel = compute_velocity_from_kinematics(padded, dt).astype(np.float32)
    qvel[:, 6:] = np.clip(qvel[:, 6:], -max_qvel, max_qvel)
    return clip.replace(velocity=qvel[:, :3], angular_velocity=qvel[:, 3:6], joints_velocity=qvel[:, 6:])


def process_clip_to_train(stac_path: str, mjcf_path: str = "./assets/rodent.xml", scale_factor: float = 0.9, start_step: int = 0,
                          clip_length: int = 250, max_qvel: float = 20.0, dt: float = 0.02, device="cuda:0",
                          _lib_path=None) -> ReferenceClip:
    """mjx_preprocess.py:44-91: STAC pickle -> clip on the model rescaled by `scale_factor` (positions and sizes)."""
    with open(stac_path, "rb") as f:
        d = pickle.load(f)
    mocap_qpos = np.asarray(d["qpos"])[start_step:start_step + clip_length]
    model = mjcf.load_xml(mjcf_path, rescale=(scale_factor, scale_factor))
    return process_clip(mocap_qpos, model, max_qvel=max_qvel, dt=dt, device=device, _lib_path=_lib_path)


# ---- storage --------------------------------------------------------------------------------------------------------
def save_reference_clip(filename: str, clip: ReferenceClip) -> None:
    """Pickle, as brax_rodent_run_ppo.py:64-73 stores it (a plain dict of arrays so that it loads without this package)."""
    with open(filename, "wb") as f:
        pickle.dump({k: (None if v is None else np.asarray(v)) for k, v in dataclasses.asdict(clip).items()}, f)


def load_reference_clip(filename: str) -> ReferenceClip:
    """Accepts this package's pickles and any object with the ReferenceClip attributes (brax_rodent_run_ppo.py:75-77)."""
    with open(filename, "rb") as f:
        d = pickle.load(f)
    names = [f.name for f in dataclasses.fields(ReferenceClip)]
    get = (lambda k: d.get(k)) if isinstance(d, dict) else (lambda k: getattr(d, k, None))
    return ReferenceClip(**{k: (None if get(k) is None else np.asarray(get(k))) for k in names})


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:  # not in this image; the pickle format above always works
        raise ImportError("HDF5 clip files need h5py (not installed); use save_reference_clip / load_reference_clip") from e


def save_reference_clip_to_h5(filename: str, clip_names: Union[List[str], str], reference_clip: ReferenceClip) -> None:
    """mjx_preprocess.py:225-252: `<clip>/<attr>` datasets; a list of names saves `value[i]` per clip."""
    assert isinstance(clip_names, (str, list))
    h5py = _h5py()
    with h5py.File(filename, "w") as hf:
        for attr, value in dataclasses.asdict(reference_clip).items():
            if value is None:
                continue
            if isinstance(clip_names, str):
                hf.create_dataset(f"{clip_names}/{attr}", data=value)
            else:
                for i, name in enumerate(clip_names):
                    hf.create_dataset(f"{name}/{attr}", data=value[i])


def load_reference_clip_from_h5(filename: str, clip_names: Union[List[str], str]) -> ReferenceClip:
    """mjx_preprocess.py:255-283: features stacked over the given clip order (leading clip axis, also for one clip)."""
    assert isinstance(clip_names, (str, list))
    h5py = _h5py()
    names = [clip_names] if isinstance(clip_names, str) else clip_names
    agg: Dict[str, list] = {}
    with h5py.File(filename, "r") as hf:
        for name in names:
            for f in dataclasses.fields(ReferenceClip):
                if f"{name}/{f.name}" in hf:
                    agg.setdefault(f.name, []).append(hf[f"{name}/{f.name}"][:])
    return ReferenceClip(**{k: np.stack(v) for k, v in agg.items()})
