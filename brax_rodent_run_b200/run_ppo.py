"""Training entry point -- the B200 counterpart of brax_rodent_run_ppo.py (same config keys, same callbacks).

    python -m brax_rodent_run_b200.run_ppo [--config readme|script] [--num-timesteps N] [--clip clips/84.p]
    torchrun --nproc-per-node 8 -m brax_rodent_run_b200.run_ppo          # envs sharded per GPU, NCCL grad all-reduce

`--config script` = the dict in brax_rodent_run_ppo.py:39-55 (1024 envs/GPU, episode 150, lr 5e-5, terminate=True, CG 8/8);
`--config readme` = readme.md:17-31 (2048 envs, episode 1000, 10M steps, batch 512, lr 3e-4, terminate=False) which is the
one BASELINE.json configs[2] names.  Without a clip file a synthetic straight-line track is used.
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import time
import uuid

import numpy as np
import torch
import torch.distributed as dist

from . import get_environment
from .ppo import PPOConfig, train

CONFIGS = {
    "script": dict(num_envs=1024, num_timesteps=500_000_000, eval_every=5_000_000, episode_length=150, batch_size=1024,
                   learning_rate=5e-5, terminate_when_unhealthy=True, solver="cg", iterations=8, ls_iterations=8, vision=False),
    "readme": dict(num_envs=2048, num_timesteps=10_000_000, eval_every=5_000_000, episode_length=1000, batch_size=512,
                   learning_rate=3e-4, terminate_when_unhealthy=False, solver="cg", iterations=8, ls_iterations=8, vision=False),
}


def load_track(path):
    if path and os.path.exists(path):
        with open(path, "rb") as f:
            clip = pickle.load(f)
        pos = clip["position"] if isinstance(clip, dict) else getattr(clip, "position", clip)  # preprocess.save_reference_clip
        return np.asarray(pos, np.float32).reshape(-1, 3)
    return np.stack([0.002 * np.arange(250), np.zeros(250), np.full(250, 0.055)], 1).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="readme", choices=sorted(CONFIGS))
    ap.add_argument("--num-timesteps", type=int)
    ap.add_argument("--num-envs", type=int)
    ap.add_argument("--model", default="rodent_new")
    ap.add_argument("--clip", default="clips/84.p")
    ap.add_argument("--out", default="./model_checkpoints")
    ap.add_argument("--eval-every", type=int, help="env steps between evaluations (config default: 5M)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tc-learner", type=int, choices=(0, 1), help="tensor-core learner off / on (default: on)")
    a = ap.parse_args()
    config = dict(CONFIGS[a.config], env_name="rodent", algo_name="ppo", task_name="run")
    if a.num_timesteps:
        config["num_timesteps"] = a.num_timesteps
    if a.num_envs:
        config["num_envs"] = a.num_envs
    if a.eval_every:
        config["eval_every"] = a.eval_every
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    torch.backends.cuda.matmul.allow_tf32 = True
    track = load_track(a.clip)
    mk = lambda n: get_environment(config["env_name"], track_pos=track, terminate_when_unhealthy=config["terminate_when_unhealthy"],
                                   solver=config["solver"], iterations=config["iterations"], ls_iterations=config["ls_iterations"],
                                   vision=config["vision"], num_envs=n, device=dev, model=a.model, kinematics_outputs=False)
    env, eval_env = mk(config["num_envs"]), mk(128)
    cfg = PPOConfig(num_timesteps=config["num_timesteps"], num_evals=max(1, int(config["num_timesteps"] / config["eval_every"])),
                    episode_length=config["episode_length"], num_envs=config["num_envs"], batch_size=config["batch_size"],
                    learning_rate=config["learning_rate"], unroll_length=10, num_minibatches=64, num_updates_per_batch=8,
                    discounting=0.97, entropy_cost=1e-3, reward_scaling=1.0, normalize_observations=True, seed=a.seed,
                    tc_learner=None if a.tc_learner is None else bool(a.tc_learner))
    run_dir = os.path.join(a.out, str(uuid.uuid4()))

    def progress(num_steps, metrics):
        print(json.dumps({"num_steps": num_steps, **{k: round(v, 5) for k, v in metrics.items()}}), flush=True)

    def policy_params_fn(num_steps, make_policy, params):
        os.makedirs(run_dir, exist_ok=True)
        with open(os.path.join(run_dir, str(num_steps)), "wb") as f:
            pickle.dump(params, f)

    t0 = time.time()
    _, agent, metrics = train(env, cfg, progress_fn=progress, policy_params_fn=policy_params_fn, eval_env=eval_env)
    if rank == 0:
        torch.save(agent.state_dict(), os.path.join(run_dir, "final.pt")) if os.path.isdir(run_dir) else None
        print(json.dumps({"done": True, "wall_s": time.time() - t0, **metrics}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
