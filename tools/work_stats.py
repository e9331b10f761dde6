"""Distribution of per-environment work cycles (rendezvous waits excluded) by warp slot within the CTA, and how much of a
CTA pass is lost to its slowest environment.   python tools/work_stats.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200.env import Rodent
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
B = 4096
env = Rodent(track, num_envs=B, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8, kinematics_outputs=False, balance=True).wrap_for_training(1000)
ctas, wpb, passes = env._geometry
print("geometry", ctas, wpb, passes)
env._env_order = lambda work: None   # the kernel's own contiguous split, but keep the work output
s = env.reset(0)
for i in range(30):
    s = env.step(s, torch.rand((B, env.action_size), device="cuda:0") * 2 - 1)
torch.cuda.synchronize()
w = s.info["work"].cpu().numpy()
e = np.arange(B)
cta = np.minimum((e * ctas + ctas - 1) // B, ctas - 1)
beg = (np.arange(ctas + 1) * B) // ctas
cta = np.searchsorted(beg, e, side="right") - 1
off = e - beg[cta]
ps, warp = off // wpb, off % wpb
print("work cycles/env-step: mean %.3g  std %.3g  min %.3g  p10 %.3g p90 %.3g max %.3g" %
      (w.mean(), w.std(), w.min(), np.percentile(w, 10), np.percentile(w, 90), w.max()))
for k in range(wpb):
    print("warp %2d (smsp %d): mean %.4g  n %d" % (k, k % 4, w[warp == k].mean(), (warp == k).sum()))
for p in range(passes):
    print("pass", p, "mean %.4g" % w[ps == p].mean(), "n", (ps == p).sum())
tot_max, tot_mean = np.zeros(ctas), np.zeros(ctas)
for c in range(ctas):
    for p in range(passes):
        g = w[(cta == c) & (ps == p)]
        if len(g):
            tot_max[c] += g.max(); tot_mean[c] += g.mean()
print("per CTA: sum over passes of max-env work: mean %.4g max %.4g ; of mean-env work: mean %.4g" % (tot_max.mean(), tot_max.max(), tot_mean.mean()))
print("=> a CTA's passes cost its slowest env: %.3f x the mean env (env-step granularity); slowest CTA / mean CTA %.3f" %
      (tot_max.mean() / tot_mean.mean(), tot_max.max() / tot_max.mean()))
print("kernel time lower bound from slowest CTA at %.3g cycles = %.3f ms @1.965GHz" % (tot_max.max(), tot_max.max() / 1.965e6))
