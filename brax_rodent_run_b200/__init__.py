"""brax_rodent_run_b200 -- B200-native batched rodent physics step + run-task env (drop-in for the hot path of
talmolab/Brax-Rodent-Run: Rodent_Env_Brax.py + the mjx.step / brax wrappers / GAE beneath it)."""
from .env import Rodent, State, PipelineState, System, load_model  # noqa: F401

_ENVS = {"rodent": Rodent}


def register_environment(name: str, cls) -> None:
    """brax.envs.register_environment (brax_rodent_run_ppo.py:57)."""
    _ENVS[name] = cls


def get_environment(name: str, **kwargs):
    """brax.envs.get_environment (brax_rodent_run_ppo.py:82-90)."""
    return _ENVS[name](**kwargs)
