"""Per-substep CG iteration counts of every environment over consecutive substeps (n_frames = 1 steps, constant action per
10 substeps as in the env step): is the per-substep solver cost periodic / predictable?  -> gpurun_out/niter_seq.npy"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200.env import Rodent
from brax_rodent_run_b200 import _lib
import ctypes
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
B = 4096
env10 = Rodent(track, num_envs=B, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8, kinematics_outputs=False)
env1 = Rodent(track, num_envs=B, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8, kinematics_outputs=False, n_frames=1)
s = env10.reset(0)
for i in range(30):
    s = env10.step(s, torch.rand((B, 30), device="cuda:0") * 2 - 1)
seq = []
niter = torch.zeros(B, dtype=torch.int32, device="cuda:0")
for step in range(6):
    a = torch.rand((B, 30), device="cuda:0") * 2 - 1
    for sub in range(10):
        buf, t = env1._out_buffers()
        ps = s.pipeline_state
        buf.in_qpos, buf.in_qvel, buf.in_act = ps.qpos.data_ptr(), ps.qvel.data_ptr(), ps.act.data_ptr()
        buf.in_qacc_warmstart, buf.in_time = ps.qacc_warmstart.data_ptr(), ps.time.data_ptr()
        buf.in_cur_frame = s.info["cur_frame"].data_ptr()
        buf.solver_niter = niter.data_ptr()
        _lib.check(env1._L, env1._L.rr_env_step(env1._env, ctypes.byref(buf), ctypes.c_void_p(a.data_ptr()), 1, env1._stream()))
        s = env1._make_state(t, a, {"cur_frame": t["cur_frame"]})
        seq.append(niter.cpu().numpy().copy())
seq = np.array(seq)  # [60, B]
np.save("gpurun_out/niter_seq.npy", seq)
print("hist", np.bincount(seq.ravel(), minlength=9) / seq.size)
print("lag-1 autocorr", np.mean([np.corrcoef(seq[t], seq[t + 1])[0, 1] for t in range(59)]))
print("lag-2 autocorr", np.mean([np.corrcoef(seq[t], seq[t + 2])[0, 1] for t in range(58)]))
for e in range(12):
    print("env", e, "".join(str(x) for x in seq[:40, e]))
