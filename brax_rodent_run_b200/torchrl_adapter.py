"""TorchRL-style view of the batched Rodent env (the role of `BraxWrapper(Rodent, ...)` in torch_utils.py:47-53).

The reference bridges the JAX env into TorchRL through dlpack (27.9 ms per step in torchrl_explore.ipynb:457); here the env already
yields torch CUDA tensors, so the adapter only renames fields into TorchRL's step convention:

    td = env.reset()                      -> {"observation" [B, obs], "done" [B, 1] bool, "terminated" [B, 1] bool}
    td = env.step({"action": a, ...})     -> input keys + {"next": {"observation", "reward" [B, 1], "done", "terminated", "truncated"}}
    td = env.step_mdp(td)                 -> td["next"] promoted to the root (what TorchRL's step_mdp does between steps)

`torchrl` / `tensordict` are not installable in this image, so plain dicts are returned; `to_tensordict()` wraps them when
tensordict is importable (batch_size = [num_envs]).  Key names and shapes follow torchrl.envs.BraxWrapper's specs
(torchrl_explore.ipynb:40-160): observation [obs_dim] float32, action [nu] in [-1, 1], reward [1], done [1] bool.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .env import Rodent, State


class RodentTorchRLEnv:
    def __init__(self, env: Rodent, episode_length: Optional[int] = 1000, seed: int = 0):
        self.env = env.wrap_for_training(episode_length) if episode_length else env
        self.batch_size = torch.Size([env.num_envs])
        self.device = env.device
        self._seed = seed
        self._state: Optional[State] = None

    # ---- specs (shapes per environment, as BraxWrapper reports them) ------------------------------------------------
    @property
    def observation_spec(self) -> Dict[str, tuple]:
        return {"observation": (self.env.observation_size,)}

    @property
    def action_spec(self) -> Dict[str, object]:
        return {"shape": (self.env.action_size,), "low": -1.0, "high": 1.0, "dtype": torch.float32}

    @property
    def reward_spec(self) -> Dict[str, tuple]:
        return {"reward": (1,)}

    def set_seed(self, seed: int) -> None:
        self._seed = int(seed)

    # ---- stepping -----------------------------------------------------------------------------------------------------
    def _view(self, st: State, with_reward: bool) -> Dict[str, torch.Tensor]:
        done = st.done.bool().unsqueeze(-1)
        out = {"observation": st.obs, "done": done}
        trunc = st.info.get("truncation")
        truncated = trunc.bool().unsqueeze(-1) if trunc is not None else torch.zeros_like(done)
        out["terminated"] = done & ~truncated
        out["truncated"] = truncated
        if with_reward:
            out["reward"] = st.reward.unsqueeze(-1)
        return out

    def reset(self, tensordict: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        self._state = self.env.reset(self._seed)
        self._seed += 1
        out = self._view(self._state, with_reward=False)
        out.pop("truncated")
        return out

    def step(self, tensordict: dict) -> dict:
        if self._state is None:
            raise RuntimeError("call reset() before step()")
        action = tensordict["action"]
        if tuple(action.shape) != (self.env.num_envs, self.env.action_size):
            raise ValueError(f"action must be [{self.env.num_envs}, {self.env.action_size}], got {tuple(action.shape)}")
        self._state = self.env.step(self._state, action)   # finished episodes restart from the cached first state (auto-reset)
        out = dict(tensordict)
        out["next"] = self._view(self._state, with_reward=True)
        return out

    @staticmethod
    def step_mdp(tensordict: dict) -> dict:
        nxt = dict(tensordict["next"])
        nxt.pop("reward", None)
        return nxt

    def rollout(self, max_steps: int, policy=None) -> Dict[str, torch.Tensor]:
        """TorchRL's env.rollout: time-stacked [B, T, ...] dict with the same nesting."""
        td = self.reset()
        steps = []
        for _ in range(max_steps):
            if policy is None:
                td["action"] = torch.rand((self.env.num_envs, self.env.action_size), device=self.device) * 2 - 1
            else:
                td = policy(td)
            td = self.step(td)
            steps.append(td)
            td = self.step_mdp(td)
        stack = lambda key_fn: torch.stack([key_fn(s) for s in steps], dim=1)
        out = {k: stack(lambda s, k=k: s[k]) for k in steps[0] if k != "next"}
        out["next"] = {k: stack(lambda s, k=k: s["next"][k]) for k in steps[0]["next"]}
        return out

    def to_tensordict(self, d: dict):
        from tensordict import TensorDict  # optional dependency (absent in this image)
        return TensorDict(d, batch_size=self.batch_size)
