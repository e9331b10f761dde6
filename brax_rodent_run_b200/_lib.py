"""ctypes binding of the C ABI in include/rr_b200.h (csrc/librr_b200.so).

The product path is the CUDA library only: `load()` raises if it has not been built -- there is no CPU
fallback.  Tests may pass an explicit `path` (the fiber emulator under tests/emu) to exercise the same host
logic and kernel text on CPU tensors.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librr_b200.so")

RR_OK, RR_EINVAL, RR_ENOTIMPL, RR_ECUDA = 0, 1, 2, 3

c_f = ctypes.POINTER(ctypes.c_float)
c_i = ctypes.POINTER(ctypes.c_int32)
c_d = ctypes.POINTER(ctypes.c_double)
vp = ctypes.c_void_p


class RRDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "ncon", "nlimit", "nefc", "nM", "nroot", "obs_dim",
                 "smem_bytes", "debug_stride")] + [("timestep", ctypes.c_float)]


_BUF_FIELDS = (
    "qpos", "qvel", "act", "qacc_warmstart", "time", "cur_frame",
    "in_qpos", "in_qvel", "in_act", "in_qacc_warmstart", "in_time", "in_cur_frame", "in_done", "in_steps",
    "obs", "reward", "done", "metrics", "steps", "truncation",
    "first_qpos", "first_qvel", "first_act", "first_qacc_warmstart", "first_time", "first_obs",
    "xpos", "xquat", "subtree_com", "qfrc_actuator", "cinert", "cvel", "contact_dist", "qacc", "solver_niter",
    "work", "env_order", "contact_pos", "contact_frame",
)


class RRBuffers(ctypes.Structure):
    """rr_buffers: every member is a raw device pointer (void* here; the header has the element types)."""
    _fields_ = [(n, vp) for n in _BUF_FIELDS]


class RRPpoLossArgs(ctypes.Structure):
    """rr_ppo_loss_args (include/rr_b200.h)."""
    _fields_ = ([(n, ctypes.c_int32) for n in ("T", "B", "A")] +
                [(n, vp) for n in ("logits", "baseline", "bootstrap", "raw_action", "old_log_prob", "reward", "discount",
                                   "truncation", "noise")] +
                [(n, ctypes.c_float) for n in ("reward_scaling", "discounting", "gae_lambda", "clipping_epsilon", "entropy_cost")] +
                [("normalize_advantage", ctypes.c_int32)] +
                [(n, vp) for n in ("scratch", "adv_partial", "loss_partial", "grad_logits", "grad_baseline")])


class RRTcProblem(ctypes.Structure):
    """rr_tc_problem (include/rr_b200.h): one GEMM of a grouped tensor-core launch."""
    _fields_ = ([(n, vp) for n in ("a", "b", "d", "bias", "aux_in", "aux_out", "ones_out")] +
                [(n, ctypes.c_int32) for n in ("m", "n", "k", "lda", "ldb", "ldd", "ldaux", "a_mn", "b_mn", "epi", "b_ones",
                                               "bn", "tile_start", "tiles_n")] +
                [("reserved", ctypes.c_int32 * 4)])


class RRGatherItem(ctypes.Structure):
    """rr_gather_item (include/rr_b200.h)."""
    _fields_ = [("src", vp), ("dst", vp)] + [(n, ctypes.c_int32) for n in ("outer", "src_rows", "inner", "dst_pitch")]


RR_POLICY_MAX_LAYERS = 8


class RRPolicyArgs(ctypes.Structure):
    """rr_policy_args (include/rr_b200.h)."""
    _fields_ = ([(n, vp) for n in ("obs", "mean", "std")] + [("w", vp * RR_POLICY_MAX_LAYERS), ("b", vp * RR_POLICY_MAX_LAYERS)] +
                [(n, vp) for n in ("eps", "action", "raw_action", "log_prob")] +
                [(n, ctypes.c_int32) for n in ("B", "obs_dim", "in0", "nlayers", "A")] + [("reserved", ctypes.c_int32 * 3)])


_libs = {}


def load(path: Optional[str] = None):
    """Load (once) and type the shared library.  `path=None` is the CUDA product library."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: the CUDA extension has not been built (run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `python -m brax_rodent_run_b200.build`).  There is no CPU fallback.")
    L = ctypes.CDLL(path)
    L.rr_last_error.restype = ctypes.c_char_p
    L.rr_model_create.argtypes = [c_i, ctypes.c_int32, c_i, ctypes.c_int32, c_d, ctypes.c_int32, ctypes.POINTER(vp)]
    L.rr_model_destroy.argtypes = [vp]
    L.rr_model_destroy.restype = None
    L.rr_model_set_solver.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
    L.rr_model_dims.argtypes = [vp, ctypes.POINTER(RRDims)]
    L.rr_env_create.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(vp)]
    L.rr_env_destroy.argtypes = [vp]
    L.rr_env_destroy.restype = None
    L.rr_env_set_task.argtypes = [vp, c_f, ctypes.c_int32, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                  ctypes.c_int32]
    L.rr_env_set_wrappers.argtypes = [vp, ctypes.c_int32]
    L.rr_env_geometry.argtypes = [vp, c_i, c_i, c_i]
    L.rr_env_init.argtypes = [vp, ctypes.POINTER(RRBuffers), vp]
    L.rr_env_step.argtypes = [vp, ctypes.POINTER(RRBuffers), vp, ctypes.c_int32, vp]
    L.rr_env_step_host.argtypes = [vp, ctypes.POINTER(RRBuffers), vp, ctypes.c_int32, vp, vp, vp, vp]
    L.rr_gae.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_float, vp, vp, vp]
    L.rr_ppo_loss_blocks.argtypes = [ctypes.c_int32, ctypes.c_int32, c_i, c_i]
    L.rr_ppo_loss.argtypes = [ctypes.POINTER(RRPpoLossArgs), vp]
    L.rr_debug_field.argtypes = [vp, ctypes.c_char_p, c_i, c_i]
    L.rr_env_set_debug.argtypes = [vp, vp]
    L.rr_env_set_profile.argtypes = [vp, vp]
    L.rr_prof_count.restype = ctypes.c_int
    L.rr_prof_name.argtypes = [ctypes.c_int32]
    L.rr_prof_name.restype = ctypes.c_char_p
    L.rr_measure_fp32_peak.argtypes = [ctypes.POINTER(ctypes.c_double), vp]
    L.rr_measure_fp32_peak.restype = ctypes.c_int
    L.rr_launch_count.restype = ctypes.c_longlong
    L.rr_adam_step.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int64] + [ctypes.c_float] * 4 + [vp]
    L.rr_adam_step.restype = ctypes.c_int
    L.rr_adam_step_sum.argtypes = [vp, vp, vp, ctypes.c_int32, vp, vp, vp, ctypes.c_int64] + [ctypes.c_float] * 4 + [vp]
    L.rr_adam_step_sum.restype = ctypes.c_int
    L.rr_gather_rows.argtypes = [ctypes.POINTER(RRGatherItem), ctypes.c_int32, vp, ctypes.c_int32, vp]
    L.rr_gather_rows.restype = ctypes.c_int
    L.rr_policy_act.argtypes = [ctypes.POINTER(RRPolicyArgs), vp]
    L.rr_policy_act.restype = ctypes.c_int
    L.rr_tc_record_bytes.restype = ctypes.c_int32
    L.rr_tc_plan.argtypes = [ctypes.POINTER(RRTcProblem), ctypes.c_int32, c_i, c_i, vp]
    L.rr_tc_plan.restype = ctypes.c_int
    L.rr_tc_launch.argtypes = [vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp]
    L.rr_tc_launch.restype = ctypes.c_int
    for name in ("rr_model_create", "rr_model_set_solver", "rr_model_dims", "rr_env_create", "rr_env_set_task",
                 "rr_env_set_wrappers", "rr_env_geometry", "rr_env_init", "rr_env_step", "rr_env_step_host", "rr_gae", "rr_debug_field",
                 "rr_env_set_debug", "rr_env_set_profile", "rr_ppo_loss", "rr_ppo_loss_blocks"):
        getattr(L, name).restype = ctypes.c_int
    _libs[path] = L
    return L


def check(L, rc: int) -> None:
    """Map a status code to the exception the reference raises in the same situation (include/rr_b200.h)."""
    if rc == RR_OK:
        return
    msg = (L.rr_last_error() or b"").decode()
    if rc == RR_EINVAL:
        raise ValueError(msg)
    if rc == RR_ENOTIMPL:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
