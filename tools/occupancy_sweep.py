"""Per-pass latency of rr_step_kernel vs environments per CTA (RR_WPB developer knob): one full pass (148 * wpb envs)."""
import os, subprocess, sys, json
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from brax_rodent_run_b200.env import Rodent
wpb = int(os.environ["RR_WPB"]); B = 148 * wpb * int(os.environ.get("PASSES", "1"))
track = np.stack([0.002*np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=B, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8, kinematics_outputs=False, balance=False).wrap_for_training(1000)
s = env.reset(0)
acts = torch.rand((40, B, 30), device="cuda:0")*2-1
for i in range(10): s = env.step(s, acts[i])
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(30): s = env.step(s, acts[10+i])
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)/30
print("wpb %2d envs %5d: %.3f ms/step  -> %.0f env-steps/s" % (wpb, B, ms, B/ms*1e3))
'''
for wpb in (1, 2, 4, 6, 8, 10):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RR_WPB=str(wpb)))
