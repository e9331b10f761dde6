"""Per-phase clock64 breakdown of rr_step_kernel (rr_env_set_profile): where one warp's cycles go.
    python tools/phase_profile.py [--envs 4096] [--steps 20] [--iterations 8] [--ls-iterations 8]
Cycles are per-warp elapsed (latency, including time the warp sat unscheduled), averaged over envs and steps."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200 import _lib  # noqa: E402
from brax_rodent_run_b200.env import Rodent  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--iterations", type=int, default=8)
ap.add_argument("--ls-iterations", type=int, default=8)
ap.add_argument("--model", default="rodent_0")
a = ap.parse_args()
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=a.envs, device="cuda:0", model=a.model, iterations=a.iterations, ls_iterations=a.ls_iterations,
             kinematics_outputs=False).wrap_for_training(1000)
L = env._L
n = L.rr_prof_count()
s = env.reset(0)
for i in range(5):
    s = env.step(s, torch.rand((a.envs, env.action_size), device="cuda:0") * 2 - 1)
prof = torch.zeros((a.envs, n), dtype=torch.int64, device="cuda:0")
niter = []
_lib.check(L, L.rr_env_set_profile(env._env, ctypes.c_void_p(prof.data_ptr())))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(a.steps):
    s = env.step(s, torch.rand((a.envs, env.action_size), device="cuda:0") * 2 - 1)
t1.record()
torch.cuda.synchronize()
L.rr_env_set_profile(env._env, None)
p = prof.double().mean(0).cpu().numpy() / a.steps
tot = p.sum()
rows = {L.rr_prof_name(i).decode(): float(p[i]) for i in range(n)}
print(json.dumps({"envs": a.envs, "steps": a.steps, "iterations": a.iterations, "ls_iterations": a.ls_iterations,
                  "ms_per_step": t0.elapsed_time(t1) / a.steps, "cycles_per_env_step": tot}))
for k, v in rows.items():
    print(f"{k:20s} {v:12.0f} cycles/env-step  {100 * v / tot:5.1f}%")
