"""Distribution of per-environment work cycles (rendezvous waits excluded) by warp slot within the CTA."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200.env import Rodent
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
B = 4096
env = Rodent(track, num_envs=B, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8, kinematics_outputs=False, balance=True).wrap_for_training(1000)
env._balance_sort = False
ctas, wpb, passes = env._geometry
print("geometry", ctas, wpb, passes)
env._env_order = lambda work: None   # identity order, but keep the work output
s = env.reset(0)
for i in range(30):
    s = env.step(s, torch.rand((B, env.action_size), device="cuda:0") * 2 - 1)
torch.cuda.synchronize()
w = s.info["work"].cpu().numpy()
slot = np.arange(B)
warp = (slot % (ctas * wpb)) % wpb
cta = (slot % (ctas * wpb)) // wpb
ps = slot // (ctas * wpb)
print("work cycles: mean %.3g  std %.3g  min %.3g  max %.3g" % (w.mean(), w.std(), w.min(), w.max()))
for k in range(wpb):
    print("warp %2d (smsp %d): mean %.4g" % (k, k % 4, w[warp == k].mean()))
for p in range(passes):
    print("pass", p, "mean %.4g" % w[ps == p].mean(), "n", (ps == p).sum())
# within-CTA spread: max/mean per (cta, pass)
ratios = []
for p in range(passes):
    for c in range(ctas):
        g = w[(ps == p) & (cta == c)]
        if len(g) == wpb:
            ratios.append(g.max() / g.mean())
print("within-CTA max/mean: mean %.3f  p90 %.3f" % (np.mean(ratios), np.percentile(ratios, 90)))
print("per-env correlation with contact count unavailable; done frac", float(s.done.mean()))
