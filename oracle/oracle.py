"""ctypes front-end of the CPU oracle (oracle/rr_oracle.c) + numpy restatement of the run-task env.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.

PARITY UNPINNED (see rr_oracle.c): mujoco-mjx 3.1.x / brax 0.10.x are not installable here.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force: bool = False) -> None:
    """Compile librr_oracle_{f32,f64}.so with gcc (no-op when up to date)."""
    src = os.path.join(_HERE, "rr_oracle.c")
    outs = [os.path.join(_HERE, f"librr_oracle_{p}.so") for p in ("f32", "f64", "cnt")]
    hdr = os.path.join(_HERE, "..", "include", "rr_model_fields.h")
    newest = max(os.path.getmtime(p) for p in (src, hdr, os.path.join(_HERE, "rr_oracle_count.cpp"), os.path.join(_HERE, "rr_opcount.hpp")))
    if not force and all(os.path.exists(o) and os.path.getmtime(o) >= newest for o in outs):
        return
    subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)


def _lib(precision: str):
    if precision not in _LIBS:
        path = os.path.join(_HERE, f"librr_oracle_{precision}.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        vp, ip, dp, ci = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double), ctypes.c_int
        L.rro_create.restype = vp
        L.rro_create.argtypes = [ip, ip, ci, dp, ci]
        L.rro_destroy.argtypes = [vp]
        L.rro_set_options.argtypes = [vp, ci, ci, ci]
        L.rro_forward.argtypes = [vp]
        L.rro_step.argtypes = [vp]
        L.rro_step_n.argtypes = [vp, ci]
        L.rro_get.restype = ci
        L.rro_get.argtypes = [vp, ctypes.c_char_p, dp, ci]
        L.rro_set.restype = ci
        L.rro_set.argtypes = [vp, ctypes.c_char_p, dp, ci]
        L.rro_scalar.restype = ctypes.c_double
        L.rro_scalar.argtypes = [vp, ctypes.c_char_p]
        L.rro_set_time.argtypes = [vp, ctypes.c_double]
        if precision == "cnt":
            L.rro_ops_total.restype = ctypes.c_longlong
            L.rro_ops_useful.restype = ctypes.c_longlong
        _LIBS[precision] = L
    return _LIBS[precision]


class Oracle:
    """One environment's mjx.Data + the mjx.forward / mjx.step restatement."""

    def __init__(self, blob, precision: str = "f32"):
        dir_, idata, fdata = blob
        self._L = _lib(precision)
        self._dir = np.ascontiguousarray(dir_, np.int32)
        self._i = np.ascontiguousarray(idata, np.int32)
        self._f = np.ascontiguousarray(fdata, np.float64)
        ip, dp = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
        self._h = self._L.rro_create(self._dir.ctypes.data_as(ip), self._i.ctypes.data_as(ip), self._i.size,
                                     self._f.ctypes.data_as(dp), self._f.size)
        self.precision = precision

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.rro_destroy(self._h)
            self._h = None

    def set_options(self, solver: int, iterations: int, ls_iterations: int):
        self._L.rro_set_options(self._h, solver, iterations, ls_iterations)

    def get(self, name: str) -> np.ndarray:
        n = self._L.rro_get(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, np.float64)
        self._L.rro_get(self._h, name.encode(), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n)
        return out

    def set(self, name: str, value) -> None:
        v = np.ascontiguousarray(value, np.float64).ravel()
        if self._L.rro_set(self._h, name.encode(), v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), v.size) < 0:
            raise KeyError(f"{name} (size {v.size})")

    def scalar(self, name: str) -> float:
        return float(self._L.rro_scalar(self._h, name.encode()))

    def forward(self):
        self._L.rro_forward(self._h)

    def ops(self, reset: bool = False):
        """(total, useful) floating-point operations executed so far (precision "cnt" only; rr_oracle_count.cpp)."""
        r = (int(self._L.rro_ops_total()), int(self._L.rro_ops_useful()))
        if reset:
            self._L.rro_ops_reset()
        return r

    def step(self, n: int = 1):
        self._L.rro_step_n(self._h, n)

    def init(self, qpos, qvel):
        """brax.mjx.pipeline.init: make_data, set qpos/qvel, mjx.forward (Rodent_Env_Brax.py:87)."""
        self.set("qpos", qpos)
        self.set("qvel", qvel)
        for name in ("act", "ctrl", "qacc_warmstart"):
            self.set(name, np.zeros_like(self.get(name)))
        self._L.rro_set_time(self._h, 0.0)
        self.forward()


class OracleRodentEnv:
    """numpy restatement of Rodent_Env_Brax.py:71-162 for ONE env, on top of `Oracle`.

    `reset` takes the already-drawn (start_frame, qpos noise, qvel noise) because the reference's
    jax.random streams cannot be reproduced without jax; everything after the draws follows the file."""

    def __init__(self, blob, dims, track_pos, ctrl_cost_weight=0.1, healthy_reward=1.0, terminate_when_unhealthy=True,
                 healthy_z_range=(0.03, 0.5), n_frames=10, solver=0, iterations=6, ls_iterations=6, precision="f32"):
        self.o = Oracle(blob, precision)
        self.o.set_options(solver, iterations, ls_iterations)
        self.nq, self.nv, self.nu, self.nbody = dims
        self.track_pos = np.asarray(track_pos, np.float64)
        self.ctrl_cost_weight, self.healthy_reward = ctrl_cost_weight, healthy_reward
        self.terminate, self.z_range, self.n_frames = terminate_when_unhealthy, healthy_z_range, n_frames
        self.cur_frame = 0

    def _track(self, idx):  # jax gather clamps out-of-range indices
        return self.track_pos[min(max(int(idx), 0), len(self.track_pos) - 1)]

    def _obs(self, cur_frame):  # Rodent_Env_Brax.py:138-162
        o = self.o
        qpos = o.get("qpos")
        xmat1 = o.get("xmat")[9:18].reshape(3, 3)
        local = xmat1 @ (self._track(cur_frame + 1) - qpos[:3])
        return np.concatenate([qpos, o.get("qvel"), o.get("cinert")[10:], o.get("cvel")[6:], o.get("qfrc_actuator"), local])

    def reset(self, start_frame, qpos, qvel):  # :71-96 after the random draws
        self.o.init(qpos, qvel)
        self.cur_frame = int(start_frame)
        return self._obs(self.cur_frame)

    def step(self, action):  # :98-136
        action = np.asarray(action, np.float64)
        self.o.set("ctrl", action)
        self.o.step(self.n_frames)
        old = self.cur_frame
        self.cur_frame += 1
        qpos = self.o.get("qpos")
        pos_reward = np.exp(-100.0 * np.linalg.norm(qpos[:3] - self._track(old)))
        healthy = 0.0 if (qpos[2] < self.z_range[0] or qpos[2] > self.z_range[1]) else 1.0
        hr = self.healthy_reward if self.terminate else self.healthy_reward * healthy
        ctrl_cost = self.ctrl_cost_weight * float(np.sum(action * action))
        obs = self._obs(self.cur_frame)
        reward = pos_reward + hr - ctrl_cost
        done = 1.0 - healthy if self.terminate else 0.0
        return obs, reward, done, dict(pos_reward=pos_reward, reward_quadctrl=-ctrl_cost, reward_alive=hr)
