"""Where a PPO minibatch update spends its time (README configuration): device time of the captured update graph, wall time of
an update including the host side, and the parts of a training step.
    python tools/ppo_update_timing.py [--tc-learner 0|1] [--replay-only N]   (--replay-only: N bare graph replays, for ncu)"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200.env import Rodent
from brax_rodent_run_b200 import ppo
ap = argparse.ArgumentParser()
ap.add_argument("--tc-learner", type=int, default=1)
ap.add_argument("--envs", type=int, default=2048)
ap.add_argument("--replay-only", type=int, default=0)
a = ap.parse_args()
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=a.envs, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8,
             terminate_when_unhealthy=False, kinematics_outputs=False)
cfg = ppo.PPOConfig(num_envs=a.envs, tc_learner=bool(a.tc_learner))
agent = ppo.PPO(env.wrap_for_training(cfg.episode_length), cfg)
state = env.reset(0)
state, _ = agent.training_step(state)
torch.cuda.synchronize()
if a.replay_only:
    for _ in range(a.replay_only):
        agent._graph.replay()
    torch.cuda.synchronize()
    sys.exit(0)
out = {"tc_learner": agent._use_tc}
# (a) device time of the update graph
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
e0.record()
for _ in range(n):
    agent._graph.replay()
e1.record(); torch.cuda.synchronize()
out["graph_replay_us"] = e0.elapsed_time(e1) / n * 1e3
# (b) a full update call (static idx copy, noise, replay), wall clock
data = agent._batch_static
idx = torch.arange(cfg.batch_size, device="cuda:0")
agent._batch_is_normalized = True
t0 = time.perf_counter()
for _ in range(n):
    agent._update_graphed(data, idx)
torch.cuda.synchronize()
out["update_call_us"] = (time.perf_counter() - t0) / n * 1e6
agent._batch_is_normalized = False
# (c) the parts of a training step
n_unroll = cfg.batch_size * cfg.num_minibatches // cfg.num_envs
t0 = time.perf_counter()
chunks = []
s2 = state
for _ in range(n_unroll):
    s2, d = agent.unroll(s2)
    chunks.append(d)
torch.cuda.synchronize()
out["rollout_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
state, _ = agent.training_step(state)
torch.cuda.synchronize()
out["training_step_s"] = time.perf_counter() - t0
out["updates_per_step"] = cfg.num_updates_per_batch * cfg.num_minibatches
out["other_s"] = out["training_step_s"] - out["rollout_s"] - out["updates_per_step"] * out["update_call_us"] * 1e-6
print(json.dumps(out))
