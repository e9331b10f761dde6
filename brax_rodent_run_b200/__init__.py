"""B200-native batched rodent physics step + run-task env (drop-in for Rodent_Env_Brax.py)."""
__version__ = "0.1.0"
