"""Per-launch device time of the tensor-core learner's grouped GEMM launches (README minibatch: 10 x 512 rows) next to the
cuBLAS TF32 time of the same contractions.    python tools/tc_learner_timing.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200.env import Rodent
from brax_rodent_run_b200 import ppo

torch.backends.cuda.matmul.allow_tf32 = True
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=2, device="cuda:0", model="rodent_0", iterations=1, ls_iterations=1, kinematics_outputs=False)
cfg = ppo.PPOConfig(num_envs=2, batch_size=2, num_minibatches=2, tc_learner=True, cuda_graph=False)
agent = ppo.PPO(env, cfg)
from brax_rodent_run_b200.tc_learner import TcLearner
tc = TcLearner(env._L, agent.policy, agent.value, 5120, 512, "cuda:0")
tc.x.normal_(); tc.xb.normal_(); tc.grad_logits.normal_(); tc.grad_baseline.normal_()


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


if "--prof" in sys.argv:  # per-CTA cycle counters of the kernel's pipeline roles
    from brax_rodent_run_b200.tc_gemm import TcGroup
    names = ["prod0 wait-empty", "prod0 issue", "prod0 wait-landed", "prod0 fence+arrive", "mma wait-full", "drain", "start->done", "epilogue", "mma issue x4", "commit"]
    for name, g in (("fwd0", tc.fwd_groups[0]), ("fwd1", tc.fwd_groups[1]), ("dgrad1", tc.dgrad_groups[1]), ("wgrad", tc.wgrad_group)):
        prof = torch.zeros(g.tiles, 16, dtype=torch.int64, device="cuda:0")
        gp = TcGroup(env._L, g._keep, "cuda:0", prof=prof)
        gp.launch(); torch.cuda.synchronize(); prof.zero_()
        gp.launch(); torch.cuda.synchronize()
        med = prof.double().median(0).values.tolist()
        mx = prof.double().max(0).values.tolist()
        print(name, "tiles", g.tiles, {n: (int(a), int(b)) for n, a, b in zip(names, med, mx)})
    sys.exit(0)

if "--once" in sys.argv:  # for ncu: every launch of one update, twice
    for _ in range(2):
        tc.forward(); tc.backward()
    torch.cuda.synchronize()
    sys.exit(0)

out = {}
for name, groups in (("fwd", tc.fwd_groups), ("dgrad", tc.dgrad_groups), ("wgrad", [tc.wgrad_group])):
    for i, g in enumerate(groups):
        out[f"{name}{i}"] = dict(us=round(timeit(g.launch), 1), tiles=g.tiles, smem=g.smem,
                                 mnk=" ".join(f"{p['m']}x{p['n']}x{p['k']}" for p in g._keep))
out["all_tc_launches"] = round(timeit(lambda: (tc.forward(), tc.backward())), 1)
x, w1, w2 = tc.x, agent.value[0].weight, agent.value[2].weight
h = tc.hv[0]
out["cublas_fwd_v0_5120x256x1264"] = round(timeit(lambda: torch.mm(x, w1.t())), 1)
out["cublas_fwd_5120x256x256"] = round(timeit(lambda: torch.mm(h, w2.t())), 1)
out["cublas_wgrad_v0_256x1264x5120"] = round(timeit(lambda: torch.mm(h.t(), x)), 1)
out["cublas_wgrad_256x256x5120"] = round(timeit(lambda: torch.mm(h.t(), h)), 1)
for k, v in out.items():
    print(k, json.dumps(v))
