import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from brax_rodent_run_b200.env import Rodent
from brax_rodent_run_b200 import ppo
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = Rodent(track, num_envs=512, device="cuda:0", model="rodent_0", iterations=2, ls_iterations=2, terminate_when_unhealthy=False, kinematics_outputs=False)
cfg = ppo.PPOConfig(num_envs=512, batch_size=512, num_minibatches=4, num_updates_per_batch=2)
agent = ppo.PPO(env.wrap_for_training(cfg.episode_length), cfg)
state = env.reset(0)
state, _ = agent.training_step(state)
torch.cuda.synchronize()
print("MARK second training step")
state, _ = agent.training_step(state)
torch.cuda.synchronize()
