import sys, numpy as np, torch, time
sys.path.insert(0, '.')
from brax_rodent_run_b200.env import Rodent
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=2048, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8, terminate_when_unhealthy=False, kinematics_outputs=False).wrap_for_training(1000)
s = env.reset(0)
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
acts = [torch.rand(2048, 30, device="cuda:0", generator=g) * 2 - 1 for _ in range(60)]
for a in acts[:10]: s = env.step(s, a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for a in acts[10:]: s = env.step(s, a)
e1.record(); torch.cuda.synchronize()
print("env.step 2048 envs, no kinematics outputs: ms/step", e0.elapsed_time(e1) / 50)
