"""TorchRL-style adapter (torch_utils.py:47-53's BraxWrapper role): key names, shapes, dtypes and the step / step_mdp
convention, checked against the underlying Rodent states."""
import numpy as np
import pytest
import torch

from conftest import backend_params, load_asset, synthetic_track


@pytest.mark.parametrize("backend", backend_params())
def test_adapter_follows_torchrl_step_convention(backend, make_env):
    from brax_rodent_run_b200.torchrl_adapter import RodentTorchRLEnv
    m = load_asset("rodent_0")
    B = 3 if backend == "emu" else 64
    base = make_env(backend, synthetic_track(), num_envs=B, model=m, iterations=2, ls_iterations=2,
                    n_frames=1 if backend == "emu" else 10)
    env = RodentTorchRLEnv(base, episode_length=3, seed=5)
    assert env.batch_size == torch.Size([B]) and env.observation_spec["observation"] == (base.observation_size,)
    td = env.reset()
    assert set(td) == {"observation", "done", "terminated"}
    assert td["observation"].shape == (B, base.observation_size) and td["done"].shape == (B, 1) and td["done"].dtype == torch.bool
    with pytest.raises(ValueError):
        env.step({"action": torch.zeros(B, 3)})
    dones = []
    for t in range(4):
        td["action"] = torch.zeros(B, base.action_size, device=base.device)
        td = env.step(td)
        nxt = td["next"]
        assert set(nxt) == {"observation", "reward", "done", "terminated", "truncated"}
        assert nxt["reward"].shape == (B, 1) and nxt["reward"].dtype == torch.float32
        assert torch.equal(nxt["done"], nxt["terminated"] | nxt["truncated"])
        assert torch.equal(nxt["observation"], env._state.obs)
        dones.append(nxt["done"].cpu().numpy().ravel().copy())
        td = env.step_mdp(td)
        assert "reward" not in td and "next" not in td and "observation" in td
    # episode_length = 3: every environment that did not terminate earlier is truncated at the third step
    assert dones[2].all()
    ro = env.rollout(5)
    assert ro["action"].shape == (B, 5, base.action_size) and ro["next"]["reward"].shape == (B, 5, 1)
    assert ro["next"]["observation"].shape == (B, 5, base.observation_size)
