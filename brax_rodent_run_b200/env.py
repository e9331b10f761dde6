"""Run-task rodent environment on the B200 step library -- the drop-in for Rodent_Env_Brax.py.

Mirrors the reference surface (Rodent_Env_Brax.py:19-162):
  Rodent(track_pos, forward_reward_weight, ctrl_cost_weight, healthy_reward, terminate_when_unhealthy,
         healthy_z_range, reset_noise_scale, solver, iterations, ls_iterations, vision, **kwargs)
  reset(rng) -> State, step(state, action) -> State with pipeline_state / obs / reward / done / metrics / info,
  observation_size, action_size, dt, sys, backend.
Differences that are deliberate: the environment is natively batched (`num_envs` rows; the reference gets its
batch from brax's VmapWrapper), tensors are torch CUDA tensors, and the physics is the hand-written sm_100a
kernel behind include/rr_b200.h instead of mjx.  Brax's EpisodeWrapper + AutoResetWrapper (applied by
ppo.train) are fused into the same kernel and enabled with `wrap_for_training`.
"""
from __future__ import annotations

import ctypes
import dataclasses
import os
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib, mjcf, model_blob

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
_XML_PATH = "./models/rodent_new.xml"  # Rodent_Env_Brax.py:16


def quat_to_mat(q: torch.Tensor) -> torch.Tensor:
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], dim=-1).reshape(q.shape[:-1] + (3, 3))


class System:
    """The slice of brax.System callers of the reference touch: sizes, qpos0, dt, names."""

    def __init__(self, model: mjcf.FlatModel):
        self.model = model
        self.nq, self.nv, self.nu, self.na = model.nq, model.nv, model.nu, model.na
        self.nbody, self.njnt, self.ngeom = model.nbody, model.njnt, model.ngeom
        self.qpos0 = model.qpos0.astype(np.float32)
        self.dt = float(model.timestep)
        self.link_names = list(model.names.get("body", [])[1:])

    def act_size(self) -> int:
        return self.nu

    def q_size(self) -> int:
        return self.nq

    def qd_size(self) -> int:
        return self.nv

    @property
    def num_links(self) -> int:
        return self.nbody - 1


@dataclasses.dataclass
class Transform:
    """brax.base.Transform: position + unit quaternion per link (world frame)."""
    pos: torch.Tensor
    rot: torch.Tensor


@dataclasses.dataclass
class Motion:
    """brax.base.Motion: angular + linear velocity per link, at the link origin (world frame)."""
    ang: torch.Tensor
    vel: torch.Tensor


@dataclasses.dataclass
class Contact:
    """brax.mjx State.contact (mjx.Contact re-typed by brax.mjx.pipeline with link_idx).  One row per
    static candidate contact (capsule pairs first, two each; then sphere / ellipsoid pairs); dist < 0 = penetrating."""
    dist: torch.Tensor          # [B, ncon]
    pos: torch.Tensor           # [B, ncon, 3]
    frame: torch.Tensor         # [B, ncon, 3, 3] rows: normal, tangent 1, tangent 2
    includemargin: torch.Tensor  # [ncon]
    friction: torch.Tensor      # [ncon, 5] (sliding x2, torsional, rolling x2 -- MuJoCo's contact friction layout)
    solref: torch.Tensor        # [ncon, 2]
    solimp: torch.Tensor        # [ncon, 5]
    geom1: torch.Tensor         # [ncon] int
    geom2: torch.Tensor         # [ncon] int
    link_idx: Tuple[torch.Tensor, torch.Tensor]  # body(geom) - 1 (-1 = world), as brax
    elasticity: torch.Tensor    # [ncon] zeros (brax fills zeros for mjx)


class PipelineState:
    """The fields of mjx.Data / brax.mjx.State the reference reads (Rodent_Env_Brax.py:110-162,
    brax_rodent_run_ppo.py:155).  cinert / cvel / qfrc_actuator are views reconstructed from the
    observation the kernel wrote, so no extra HBM traffic is spent on them."""

    def __init__(self, env: "Rodent", qpos, qvel, act, qacc_warmstart, time, ctrl, obs=None, xpos=None, xquat=None,
                 subtree_com=None, contact_dist=None, contact_pos=None, contact_frame=None):
        self._env = env
        self.qpos, self.qvel, self.act, self.qacc_warmstart, self.time, self.ctrl = qpos, qvel, act, qacc_warmstart, time, ctrl
        self._obs = obs
        self.xpos, self.xquat, self.subtree_com, self.contact_dist = xpos, xquat, subtree_com, contact_dist
        self._contact_pos, self._contact_frame = contact_pos, contact_frame

    # brax.mjx.State views (brax.mjx.pipeline.step) ------------------------------------------------
    @property
    def x(self) -> Transform:
        """Link transforms: x = Transform(xpos[1:], xquat[1:])."""
        if self.xquat is None:
            raise AttributeError("x needs kinematics outputs (Rodent(..., kinematics_outputs=True))")
        return Transform(pos=self.xpos[:, 1:], rot=self.xquat[:, 1:])

    @property
    def xd(self) -> Motion:
        """Link velocities at the link origins: cvel[1:] (ang, lin about the tree's subtree COM) moved by
        offset = xpos[1:] - subtree_com[root]: vel = lin - offset x ang (brax: Transform.create(pos=offset).do(Motion))."""
        if self.xquat is None:
            raise AttributeError("xd needs kinematics outputs (Rodent(..., kinematics_outputs=True))")
        cv = self.cvel[:, 1:]
        root = self._env._body_rootslot[1:]
        offset = self.xpos[:, 1:] - self.subtree_com[:, root]
        ang, lin = cv[..., :3], cv[..., 3:]
        return Motion(ang=ang, vel=lin - torch.linalg.cross(offset, ang))

    @property
    def contact(self) -> Contact:
        if self.contact_dist is None:
            raise AttributeError("contact needs kinematics outputs (Rodent(..., kinematics_outputs=True))")
        c = self._env._contact_const
        B, n = self.contact_dist.shape
        return Contact(dist=self.contact_dist, pos=self._contact_pos, frame=self._contact_frame.reshape(B, n, 3, 3),
                       includemargin=c["includemargin"], friction=c["friction"], solref=c["solref"], solimp=c["solimp"],
                       geom1=c["geom1"], geom2=c["geom2"], link_idx=c["link_idx"], elasticity=c["elasticity"])

    # brax aliases
    @property
    def q(self):
        return self.qpos

    @property
    def qd(self):
        return self.qvel

    @property
    def xmat(self):
        if self.xquat is None:
            raise AttributeError("xmat needs kinematics outputs (Rodent(..., kinematics_outputs=True))")
        return quat_to_mat(self.xquat)

    def _obs_slice(self, lo, hi, shape):
        if self._obs is None:
            raise AttributeError("this pipeline_state was produced without an observation buffer")
        B = self._obs.shape[0]
        body = self._obs[:, lo:hi].reshape((B,) + shape)
        return torch.cat([torch.zeros((B, 1) + shape[1:], dtype=body.dtype, device=body.device), body], dim=1)

    @property
    def cinert(self):
        s = self._env.sys
        lo = s.nq + s.nv
        return self._obs_slice(lo, lo + 10 * (s.nbody - 1), (s.nbody - 1, 10))

    @property
    def cvel(self):
        s = self._env.sys
        lo = s.nq + s.nv + 10 * (s.nbody - 1)
        return self._obs_slice(lo, lo + 6 * (s.nbody - 1), (s.nbody - 1, 6))

    @property
    def qfrc_actuator(self):
        s = self._env.sys
        lo = s.nq + s.nv + 16 * (s.nbody - 1)
        return self._obs[:, lo:lo + s.nv]


@dataclasses.dataclass
class State:
    """brax.envs.base.State"""
    pipeline_state: PipelineState
    obs: torch.Tensor
    reward: torch.Tensor
    done: torch.Tensor
    metrics: Dict[str, torch.Tensor] = dataclasses.field(default_factory=dict)
    info: Dict[str, Any] = dataclasses.field(default_factory=dict)

    def replace(self, **kw) -> "State":
        return dataclasses.replace(self, **kw)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def load_model(model: Union[str, mjcf.FlatModel, None] = None, xml_path: Optional[str] = None) -> mjcf.FlatModel:
    """mujoco.MjModel.from_xml_path (Rodent_Env_Brax.py:41): parse an MJCF file, or load one of the compiled
    rodent assets by name ('rodent_0', 'rodent_new', 'rodent_optimized', 'rodent_pair')."""
    if isinstance(model, mjcf.FlatModel):
        return model
    if xml_path is not None:
        return mjcf.load_xml(xml_path)
    name = model or os.path.splitext(os.path.basename(_XML_PATH))[0]
    path = os.path.join(_ASSETS, name + ".npz")
    if not os.path.exists(path):
        raise ValueError(f"XML Error: unknown model '{name}' (no compiled asset {path})")
    return mjcf.FlatModel.load(path)


class Rodent:
    """Rodent run task (Rodent_Env_Brax.py:19)."""

    def __init__(
        self,
        track_pos,
        forward_reward_weight=10,
        ctrl_cost_weight=0.1,
        healthy_reward=1.0,
        terminate_when_unhealthy=True,
        healthy_z_range=(0.03, 0.5),
        reset_noise_scale=1e-2,
        solver="cg",
        iterations: int = 6,
        ls_iterations: int = 6,
        vision=False,
        *,
        num_envs: int = 1,
        device: Union[str, torch.device] = "cuda:0",
        model: Union[str, mjcf.FlatModel, None] = None,
        xml_path: Optional[str] = None,
        kinematics_outputs: bool = True,
        balance: bool = False,
        reuse_buffers: bool = False,
        _lib_path: Optional[str] = None,
        **kwargs,
    ):
        self._L = _lib.load(_lib_path)
        self.device = torch.device(device)
        if self.device.type != "cuda" and _lib_path is None:
            raise RuntimeError("brax_rodent_run_b200 runs on CUDA devices only (there is no CPU fallback)")
        flat = load_model(model, xml_path)
        solver_id = {"cg": 0, "newton": 1}[solver.lower()]  # KeyError on anything else, as the reference (:42-45)
        self._n_frames = int(kwargs.pop("n_frames", 10))  # :53-57
        kwargs.pop("backend", None)
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        self.sys = System(flat)
        self.backend = "b200"
        self.num_envs = int(num_envs)
        self._kin = bool(kinematics_outputs)
        # reuse_buffers: step() cycles through THREE preallocated output sets instead of allocating ten tensors per call.  A
        # State stays valid until the third step() after the one that produced it (enough for `s = env.step(s, a)` loops
        # and for holding on to the previous state); anything kept longer must be cloned.  Off by default: brax States are
        # immutable, and that is what the default gives.
        self._reuse = bool(reuse_buffers)
        self._ring, self._ring_pos = [], 0

        dir_, idata, fdata = model_blob.pack(flat)
        self._blob = (np.ascontiguousarray(dir_, np.int32), np.ascontiguousarray(idata, np.int32),
                      np.ascontiguousarray(fdata, np.float64))
        h = ctypes.c_void_p()
        _lib.check(self._L, self._L.rr_model_create(
            self._blob[0].ctypes.data_as(_lib.c_i), self._blob[0].size, self._blob[1].ctypes.data_as(_lib.c_i),
            self._blob[1].size, self._blob[2].ctypes.data_as(_lib.c_d), self._blob[2].size, ctypes.byref(h)))
        self._model = h
        _lib.check(self._L, self._L.rr_model_set_solver(self._model, solver_id, int(iterations), int(ls_iterations)))
        self.dims = _lib.RRDims()
        _lib.check(self._L, self._L.rr_model_dims(self._model, ctypes.byref(self.dims)))
        e = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else 0
        _lib.check(self._L, self._L.rr_env_create(self._model, self.num_envs, dev_index, ctypes.byref(e)))
        self._env = e

        tp = track_pos.detach().cpu().numpy() if isinstance(track_pos, torch.Tensor) else np.asarray(track_pos)
        tp = np.ascontiguousarray(tp, np.float32).reshape(-1, 3)
        self._track_pos_np = tp
        self._track_pos = torch.from_numpy(tp).to(self.device)
        self._forward_reward_weight = forward_reward_weight
        self._ctrl_cost_weight = float(ctrl_cost_weight)
        self._healthy_reward = float(healthy_reward)
        self._terminate_when_unhealthy = bool(terminate_when_unhealthy)
        self._healthy_z_range = (float(healthy_z_range[0]), float(healthy_z_range[1]))
        self._reset_noise_scale = float(reset_noise_scale)
        self._vision = vision
        _lib.check(self._L, self._L.rr_env_set_task(
            self._env, tp.ctypes.data_as(_lib.c_f), tp.shape[0], self._ctrl_cost_weight, self._healthy_reward,
            self._healthy_z_range[0], self._healthy_z_range[1], int(self._terminate_when_unhealthy)))
        self._episode_length = 0
        self._qpos0 = torch.from_numpy(self.sys.qpos0).to(self.device)
        self._init_views(flat)
        # load balancing: the warps of a CTA rendezvous every substep, so environments of similar cost (last step's
        # cycle count) are grouped into the same CTA and the groups dealt to the CTAs in snake order
        g, w, p = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _lib.check(self._L, self._L.rr_env_geometry(self._env, ctypes.byref(g), ctypes.byref(w), ctypes.byref(p)))
        self._geometry = (g.value, w.value, p.value)
        self._balance = bool(balance) and self.num_envs > w.value
        self._slot_of_group = self._snake_slots() if self._balance else None

    def _init_views(self, flat: mjcf.FlatModel) -> None:
        """Constants of the brax.mjx State views (x / xd / contact)."""
        dev = self.device
        roots, slot = [], np.zeros(flat.nbody, np.int64)
        for b in range(1, flat.nbody):  # slot = order of first appearance of the tree root (as csrc/rr_model_build.h)
            r = int(flat.body_rootid[b])
            if r not in roots:
                roots.append(r)
            slot[b] = roots.index(r)
        self._body_rootslot = torch.from_numpy(slot).to(dev)
        npair = int(flat.npair)
        conadr = np.asarray(flat.pair_conadr, np.int64)
        cnt = np.diff(np.append(conadr, flat.ncon)) if npair else np.zeros(0, np.int64)
        con_pair = np.repeat(np.arange(npair), cnt)
        g1, g2 = np.asarray(flat.pair_geom1, np.int64)[con_pair], np.asarray(flat.pair_geom2, np.int64)[con_pair]
        gb = np.asarray(flat.geom_bodyid, np.int64)
        fr = np.asarray(flat.pair_friction, np.float32).reshape(npair, -1)[con_pair] if npair else np.zeros((0, 1), np.float32)
        gf1, gf2 = np.asarray(flat.geom_friction, np.float32)[g1], np.asarray(flat.geom_friction, np.float32)[g2]
        pr1, pr2 = np.asarray(flat.geom_priority)[g1], np.asarray(flat.geom_priority)[g2]
        f3 = np.where((pr1 > pr2)[:, None], gf1, np.where((pr2 > pr1)[:, None], gf2, np.maximum(gf1, gf2)))  # MuJoCo mixing rule
        friction = np.stack([fr[:, 0], fr[:, 0], f3[:, 1], f3[:, 2], f3[:, 2]], 1) if npair else np.zeros((0, 5), np.float32)
        tt = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
        self._contact_const = dict(
            includemargin=tt(np.asarray(flat.pair_includemargin, np.float32)[con_pair]), friction=tt(friction),
            solref=tt(np.asarray(flat.pair_solref, np.float32).reshape(npair, 2)[con_pair]),
            solimp=tt(np.asarray(flat.pair_solimp, np.float32).reshape(npair, 5)[con_pair]),
            geom1=tt(g1, torch.int64), geom2=tt(g2, torch.int64),
            link_idx=(tt(gb[g1] - 1, torch.int64), tt(gb[g2] - 1, torch.int64)),
            elasticity=torch.zeros(len(con_pair), device=dev))

    def _snake_slots(self) -> torch.Tensor:
        """Slot filled by the k-th most expensive environment.  Environments of similar cost share a CTA pass (the warps
        of a CTA rendezvous, so a pass costs its slowest environment); groups are dealt to the CTAs in snake order (pass 0
        CTA 0..n-1, pass 1 in reverse, ...) so that every CTA gets a similar total, and the short last pass is spread
        evenly over the CTAs (idle warps only run the barriers)."""
        ctas, wpb, passes = self._geometry
        B = self.num_envs
        last = B - ctas * wpb * (passes - 1)           # environments in the last pass
        slots = []
        for ps in range(passes):
            ctas_order = range(ctas) if ps % 2 == 0 else range(ctas - 1, -1, -1)
            for c in ctas_order:
                cap = wpb if ps < passes - 1 else (last * (c + 1)) // ctas - (last * c) // ctas
                slots += [ps * ctas * wpb + c * wpb + w for w in range(cap)]
        assert len(slots) == B
        return torch.tensor(slots, dtype=torch.long, device=self.device)

    def _env_order(self, work: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if not self._balance or work is None:
            return None
        ctas, wpb, passes = self._geometry
        idx = torch.argsort(work, descending=True).to(torch.int32)
        out = torch.full((ctas * wpb * passes,), -1, dtype=torch.int32, device=self.device)
        out[self._slot_of_group] = idx
        return out

    def __del__(self):
        L = getattr(self, "_L", None)
        if L is None:
            return
        if getattr(self, "_env", None):
            L.rr_env_destroy(self._env)
            self._env = None
        if getattr(self, "_model", None):
            L.rr_model_destroy(self._model)
            self._model = None

    # ---- brax Env properties --------------------------------------------------------------------------
    @property
    def observation_size(self) -> int:
        return int(self.dims.obs_dim)

    @property
    def action_size(self) -> int:
        return self.sys.nu

    @property
    def dt(self) -> float:
        return self.sys.dt * self._n_frames

    @property
    def n_frames(self) -> int:
        return self._n_frames

    @property
    def unwrapped(self) -> "Rodent":
        return self

    # ---- fused training wrappers ------------------------------------------------------------------------
    def wrap_for_training(self, episode_length: int = 1000, action_repeat: int = 1) -> "Rodent":
        """brax.envs.wrappers.training.wrap: Vmap (native) + EpisodeWrapper + AutoResetWrapper, fused in-kernel."""
        if action_repeat != 1:
            raise NotImplementedError("action_repeat != 1 (the reference uses 1, brax_rodent_run_ppo.py:104)")
        _lib.check(self._L, self._L.rr_env_set_wrappers(self._env, int(episode_length)))
        self._episode_length = int(episode_length)
        return self

    # ---- helpers ----------------------------------------------------------------------------------------
    def _stream(self):
        if self.device.type == "cuda":
            return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return None

    def _empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _out_buffers(self, ring: bool = False) -> Tuple[_lib.RRBuffers, Dict[str, torch.Tensor]]:
        if ring and self._reuse:
            if len(self._ring) < 3:
                self._ring.append(self._out_buffers())
                return self._ring[-1]
            self._ring_pos = (self._ring_pos + 1) % 3
            buf, t = self._ring[self._ring_pos]
            fresh = _lib.RRBuffers()
            for k, v in t.items():
                setattr(fresh, k, v.data_ptr())
            return fresh, t
        B, d = self.num_envs, self.dims
        t = dict(qpos=self._empty(B, d.nq), qvel=self._empty(B, d.nv), act=self._empty(B, d.na),
                 qacc_warmstart=self._empty(B, d.nv), time=self._empty(B), cur_frame=self._empty(B, dtype=torch.int32),
                 obs=self._empty(B, d.obs_dim), reward=self._empty(B), done=self._empty(B), metrics=self._empty(B, 3))
        if self._kin:
            t.update(xpos=self._empty(B, d.nbody, 3), xquat=self._empty(B, d.nbody, 4), subtree_com=self._empty(B, d.nroot, 3))
            if d.ncon > 0:
                t.update(contact_dist=self._empty(B, d.ncon), contact_pos=self._empty(B, d.ncon, 3),
                         contact_frame=self._empty(B, d.ncon, 9))
        if self._episode_length:
            t.update(steps=self._empty(B), truncation=self._empty(B))
        if self._balance:
            t.update(work=self._empty(B))
        buf = _lib.RRBuffers()
        for k, v in t.items():
            setattr(buf, k, v.data_ptr())
        return buf, t

    def _make_state(self, t, ctrl, info) -> State:
        ps = PipelineState(self, t["qpos"], t["qvel"], t["act"], t["qacc_warmstart"], t["time"], ctrl, obs=t["obs"],
                           xpos=t.get("xpos"), xquat=t.get("xquat"), subtree_com=t.get("subtree_com"),
                           contact_dist=t.get("contact_dist"), contact_pos=t.get("contact_pos"),
                           contact_frame=t.get("contact_frame"))
        m = t["metrics"]
        metrics = {"pos_reward": m[:, 0], "reward_quadctrl": m[:, 1], "reward_alive": m[:, 2]}
        return State(ps, t["obs"], t["reward"], t["done"], metrics, info)

    # ---- reset (Rodent_Env_Brax.py:71-96) -----------------------------------------------------------------
    def reset(self, rng) -> State:
        """`rng`: int seed or torch.Generator on the env's device.  (jax.random key streams are not reproduced;
        the distributions are: start_frame ~ randint[0, 100), qpos/qvel noise ~ U(-scale, scale).)"""
        if isinstance(rng, torch.Generator):
            gen = rng
        else:
            gen = torch.Generator(device=self.device)
            gen.manual_seed(int(rng))
        B, s = self.num_envs, self._reset_noise_scale
        start_frame = torch.randint(0, 100, (B,), generator=gen, device=self.device, dtype=torch.int32)
        noise_q = (torch.rand((B, self.sys.nq), generator=gen, device=self.device) * 2 - 1) * s
        noise_v = (torch.rand((B, self.sys.nv), generator=gen, device=self.device) * 2 - 1) * s
        return self.reset_from(start_frame, noise_q, noise_v)

    def reset_from(self, start_frame: torch.Tensor, qpos_noise: torch.Tensor, qvel_noise: torch.Tensor) -> State:
        """Everything in Rodent.reset after the random draws (:77-96)."""
        B = self.num_envs
        start_frame = start_frame.to(self.device, torch.int32).reshape(B)
        idx = start_frame.long().clamp(0, self._track_pos.shape[0] - 1)
        qpos = self._qpos0.repeat(B, 1)
        qpos[:, :3] = self._track_pos[idx]
        qpos = qpos + qpos_noise.to(self.device, torch.float32)
        qvel = qvel_noise.to(self.device, torch.float32).clone()
        return self.init_state(qpos, qvel, start_frame)

    def init_state(self, qpos: torch.Tensor, qvel: torch.Tensor, cur_frame: Optional[torch.Tensor] = None) -> State:
        """pipeline_init (:87) + _get_obs + zero reward/done/metrics (:89-96)."""
        B = self.num_envs
        buf, t = self._out_buffers()
        t["qpos"].copy_(qpos.reshape(B, -1))
        t["qvel"].copy_(qvel.reshape(B, -1))
        t["act"].zero_()
        t["qacc_warmstart"].zero_()
        t["time"].zero_()
        if cur_frame is None:
            t["cur_frame"].zero_()
        else:
            t["cur_frame"].copy_(cur_frame.reshape(B))
        _lib.check(self._L, self._L.rr_env_init(self._env, ctypes.byref(buf), self._stream()))
        info: Dict[str, Any] = {"cur_frame": t["cur_frame"]}
        ctrl = torch.zeros((B, self.sys.nu), device=self.device)
        state = self._make_state(t, ctrl, info)
        if self._episode_length:
            info["steps"], info["truncation"] = t["steps"], t["truncation"]
            info["first_pipeline_state"] = state.pipeline_state
            info["first_obs"] = state.obs
        return state

    # ---- step (Rodent_Env_Brax.py:98-136) -------------------------------------------------------------------
    def step(self, state: State, action: torch.Tensor) -> State:
        B = self.num_envs
        action = action.to(self.device, torch.float32).reshape(B, self.sys.nu).contiguous()
        buf, t = self._out_buffers(ring=True)
        ps = state.pipeline_state
        buf.in_qpos, buf.in_qvel, buf.in_act = ps.qpos.data_ptr(), ps.qvel.data_ptr(), ps.act.data_ptr()
        buf.in_qacc_warmstart, buf.in_time = ps.qacc_warmstart.data_ptr(), ps.time.data_ptr()
        buf.in_cur_frame = state.info["cur_frame"].data_ptr()
        info = dict(state.info)
        if self._episode_length:
            buf.in_done, buf.in_steps = state.done.data_ptr(), state.info["steps"].data_ptr()
            f = state.info["first_pipeline_state"]
            buf.first_qpos, buf.first_qvel, buf.first_act = f.qpos.data_ptr(), f.qvel.data_ptr(), f.act.data_ptr()
            buf.first_qacc_warmstart, buf.first_time = f.qacc_warmstart.data_ptr(), f.time.data_ptr()
            buf.first_obs = state.info["first_obs"].data_ptr()
        order = self._env_order(state.info.get("work"))
        if order is not None:
            buf.env_order = order.data_ptr()
        _lib.check(self._L, self._L.rr_env_step(self._env, ctypes.byref(buf), _ptr(action), self._n_frames, self._stream()))
        info["cur_frame"] = t["cur_frame"]
        if self._balance:
            info["work"] = t["work"]
        if self._episode_length:
            info["steps"], info["truncation"] = t["steps"], t["truncation"]
        return self._make_state(t, action, info)

    # ---- parity-test hook -------------------------------------------------------------------------------------
    def debug_forward(self, qpos, qvel, act=None, ctrl=None, qacc_warmstart=None) -> Dict[str, torch.Tensor]:
        """Run one mjx.forward and return the intermediates named in rr_debug_field (tests only)."""
        B, d = self.num_envs, self.dims
        dbg = torch.zeros((B, d.debug_stride), device=self.device)
        _lib.check(self._L, self._L.rr_env_set_debug(self._env, _ptr(dbg)))
        try:
            buf = _lib.RRBuffers()
            t = dict(qpos=qpos.to(self.device, torch.float32).reshape(B, d.nq).clone(),
                     qvel=qvel.to(self.device, torch.float32).reshape(B, d.nv).clone(),
                     act=(torch.zeros((B, d.na), device=self.device) if act is None else act.to(self.device, torch.float32).clone()),
                     qacc_warmstart=(torch.zeros((B, d.nv), device=self.device) if qacc_warmstart is None
                                     else qacc_warmstart.to(self.device, torch.float32).clone()))
            for k, v in t.items():
                setattr(buf, k, v.data_ptr())
            if ctrl is None:
                _lib.check(self._L, self._L.rr_env_init(self._env, ctypes.byref(buf), self._stream()))
            else:
                raise NotImplementedError("debug_forward with ctrl: use step()")
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
        finally:
            self._L.rr_env_set_debug(self._env, None)
        out = {"qpos": t["qpos"], "qacc_warmstart": t["qacc_warmstart"]}
        off, cnt = ctypes.c_int32(), ctypes.c_int32()
        for name in ("xpos", "xquat", "subtree_com", "cinert", "cdof", "cvel", "qM_sparse", "qLD_sparse", "qfrc_bias",
                     "qfrc_passive", "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "contact_dist", "contact_pos",
                     "contact_frame", "efc_J", "efc_D", "efc_aref", "efc_force", "qacc", "qfrc_constraint", "scalars"):
            _lib.check(self._L, self._L.rr_debug_field(self._model, name.encode(), ctypes.byref(off), ctypes.byref(cnt)))
            out[name] = dbg[:, off.value:off.value + cnt.value]
        return out
