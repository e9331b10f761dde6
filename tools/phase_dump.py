"""Per-environment phase cycle counts of ONE env step (10 substeps) -> gpurun_out/phase_dump.npz (prof [B, NPROF], names, niter)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200 import _lib
from brax_rodent_run_b200.env import Rodent
B = 4096
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=B, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8, kinematics_outputs=False).wrap_for_training(1000)
L = env._L
n = L.rr_prof_count()
s = env.reset(0)
for i in range(30):
    s = env.step(s, torch.rand((B, env.action_size), device="cuda:0") * 2 - 1)
out = []
for rep in range(4):
    prof = torch.zeros((B, n), dtype=torch.int64, device="cuda:0")
    _lib.check(L, L.rr_env_set_profile(env._env, ctypes.c_void_p(prof.data_ptr())))
    s = env.step(s, torch.rand((B, env.action_size), device="cuda:0") * 2 - 1)
    torch.cuda.synchronize()
    out.append(prof.cpu().numpy())
L.rr_env_set_profile(env._env, None)
names = [L.rr_prof_name(i).decode() for i in range(n)]
np.savez("gpurun_out/phase_dump.npz", prof=np.array(out), names=np.array(names), geometry=np.array(env._geometry))
print(names, env._geometry)
