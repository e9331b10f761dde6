import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from brax_rodent_run_b200 import get_environment
from brax_rodent_run_b200.ppo import PPO, PPOConfig, tanh_normal_log_prob
track = np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)
env = get_environment("rodent", track_pos=track, terminate_when_unhealthy=False, iterations=8, ls_iterations=8, num_envs=2048,
                      device="cuda:0", model="rodent_0", kinematics_outputs=False).wrap_for_training(1000)
cfg = PPOConfig(num_envs=2048, batch_size=512)
agent = PPO(env, cfg)
st = env.reset(0)
print("reset obs finite", torch.isfinite(st.obs).all().item(), "absmax", st.obs.abs().max().item())
for u in range(3):
    st, data = agent.unroll(st)
    for k, v in data.items():
        print(u, k, "finite", torch.isfinite(v).all().item(), "absmax", v.abs().max().item())
agent.normalizer.update(data["observation"])
print("std min/max", agent.normalizer.std.min().item(), agent.normalizer.std.max().item())
nobs = agent.normalizer.normalize(data["observation"])
print("normalized absmax", nobs.abs().max().item(), "argmax col", (nobs.abs().amax((0,1))).argmax().item())
logits = agent.policy(nobs)
lp = tanh_normal_log_prob(logits, data["raw_action"])
print("logits absmax", logits.abs().max().item(), "target lp min/max", lp.min().item(), lp.max().item(), "behaviour lp min/max", data["log_prob"].min().item(), data["log_prob"].max().item())
print("qpos z range", st.pipeline_state.qpos[:,2].min().item(), st.pipeline_state.qpos[:,2].max().item(), "qvel absmax", st.pipeline_state.qvel.abs().max().item())
