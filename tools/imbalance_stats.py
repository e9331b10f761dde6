"""How uneven is the per-environment cost that the CTA-wide rendezvous has to absorb?
    python tools/imbalance_stats.py
Uses the instrumented kernel's per-environment phase cycles for single control steps: distribution of the solver-phase
cycles across environments, step-to-step persistence (correlation), and the solver iteration histogram."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200 import _lib
from brax_rodent_run_b200.env import Rodent
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
B = 4096
env = Rodent(track, num_envs=B, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8, kinematics_outputs=False,
             balance=False).wrap_for_training(1000)
L = env._L
n = L.rr_prof_count()
names = [L.rr_prof_name(i).decode() for i in range(n)]
s = env.reset(0)
for i in range(30):
    s = env.step(s, torch.rand((B, env.action_size), device="cuda:0") * 2 - 1)
per_step = []
for k in range(4):
    prof = torch.zeros((B, n), dtype=torch.int64, device="cuda:0")
    _lib.check(L, L.rr_env_set_profile(env._env, ctypes.c_void_p(prof.data_ptr())))
    s = env.step(s, torch.rand((B, env.action_size), device="cuda:0") * 2 - 1)
    torch.cuda.synchronize()
    L.rr_env_set_profile(env._env, None)
    per_step.append(prof.double().cpu().numpy())
sol = [i for i, nm in enumerate(names) if nm.startswith("solver") or nm in ("make_constraint", "collision", "euler")]
print("solver-side buckets:", [names[i] for i in sol])
tot = [p.sum(1) for p in per_step]
sv = [p[:, sol].sum(1) for p in per_step]
for k in range(4):
    print("step %d: solver-side cycles/env-step mean %.3g std %.3g p10 %.3g p90 %.3g max %.3g" %
          (k, sv[k].mean(), sv[k].std(), np.percentile(sv[k], 10), np.percentile(sv[k], 90), sv[k].max()))
print("corr(step0, step1) %.3f   corr(step0, step3) %.3f" % (np.corrcoef(sv[0], sv[1])[0, 1], np.corrcoef(sv[0], sv[3])[0, 1]))
# grouping experiment: mean over groups of 10 of max/mean, natural order vs sorted by the previous step's cost
def grp(x, order):
    g = x[order][: (B // 10) * 10].reshape(-1, 10)
    return float((g.max(1) / g.mean(1)).mean())
nat = np.arange(B)
print("within-group-of-10 max/mean of solver-side cost: natural %.3f, sorted by previous step %.3f, sorted by itself %.3f" %
      (grp(sv[1], nat), grp(sv[1], np.argsort(sv[0])), grp(sv[1], np.argsort(sv[1]))))
print("per-bucket cycles/env-step across environments (one step): p10 / p50 / p90")
for i, nm in enumerate(names):
    x = per_step[1][:, i]
    print("  %-20s %10.0f %10.0f %10.0f" % (nm, np.percentile(x, 10), np.percentile(x, 50), np.percentile(x, 90)))
x = per_step[1].sum(1)
print("  %-20s %10.0f %10.0f %10.0f" % ("total", np.percentile(x, 10), np.percentile(x, 50), np.percentile(x, 90)))
# solver iterations of the last substep
import ctypes as C
buf, t = env._out_buffers()
nit = torch.zeros(B, dtype=torch.int32, device="cuda:0")
env2 = env
orig = env._out_buffers
def patched():
    b, tt = orig()
    b.solver_niter = nit.data_ptr()
    return b, tt
env._out_buffers = patched
s = env.step(s, torch.rand((B, env.action_size), device="cuda:0") * 2 - 1)
torch.cuda.synchronize()
print("solver iterations (last substep) histogram:", np.bincount(nit.cpu().numpy(), minlength=10).tolist())
