"""PPO train SPS on the README configuration (BASELINE configs[2]): 2048 envs, unroll 10, batch 512 x 64 minibatches, 8 epochs.
    python tools/ppo_sps.py [--train-steps 3] [--envs 2048]
Prints the split of a training step into rollout (policy + env steps) and learner (epochs x minibatches) time."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200.env import Rodent
from brax_rodent_run_b200 import ppo
ap = argparse.ArgumentParser()
ap.add_argument("--train-steps", type=int, default=3)
ap.add_argument("--envs", type=int, default=2048)
ap.add_argument("--tc-learner", type=int, default=None, help="1 / 0: tensor-core learner on / off (default: PPOConfig's)")
a = ap.parse_args()
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=a.envs, device="cuda:0", model="rodent_0", iterations=8, ls_iterations=8,
             terminate_when_unhealthy=False, kinematics_outputs=False)
cfg = ppo.PPOConfig(num_envs=a.envs) if a.tc_learner is None else ppo.PPOConfig(num_envs=a.envs, tc_learner=bool(a.tc_learner))
agent = ppo.PPO(env.wrap_for_training(cfg.episode_length), cfg)
state = env.reset(0)
state, _ = agent.training_step(state)  # warm-up (allocator, cuBLAS handles)
torch.cuda.synchronize()
# split: time the rollout part alone
n_unroll = cfg.batch_size * cfg.num_minibatches // cfg.num_envs
t0 = time.perf_counter()
s2 = state
for _ in range(n_unroll):
    s2, _d = agent.unroll(s2)
torch.cuda.synchronize()
t_roll = time.perf_counter() - t0
e0 = agent.env_steps
t0 = time.perf_counter()
for _ in range(a.train_steps):
    state, m = agent.training_step(state)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
steps = agent.env_steps - e0
print(json.dumps({"metric": "PPO train SPS", "value": steps / dt, "unit": "env-steps/s", "envs": a.envs, "train_steps": a.train_steps,
                  "env_steps_per_train_step": steps // a.train_steps, "s_per_train_step": dt / a.train_steps,
                  "tc_learner": agent._use_tc, "rollout_s": t_roll, "learner_s": dt / a.train_steps - t_roll,
                  "config": "README: batch 512, 64 minibatches, unroll 10, 8 epochs, CG 8/8, normalize obs"}))
