"""A 64-environment reset + env step (rodent_0, CG 4/4, 3 substeps, fused wrappers) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_step.py
    compute-sanitizer --tool racecheck python tools/sanitize_step.py
(one tool per gpurun call; logs kept under profiles/)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200.env import Rodent
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=64, device="cuda:0", model="rodent_0", iterations=4, ls_iterations=4, n_frames=3).wrap_for_training(1000)
s = env.reset(0)
for i in range(2):
    s = env.step(s, torch.rand((64, env.action_size), device="cuda:0") * 2 - 1)
torch.cuda.synchronize()
print("sanitize_step ok: reward mean %.4f, finite %s" % (float(s.reward.mean()), bool(torch.isfinite(s.obs).all())))
