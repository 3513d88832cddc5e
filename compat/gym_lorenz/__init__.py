"""Drop-in stand-in for the reference package `gym_lorenz` (gym_lorenz/__init__.py:1-23).

Put this directory on `sys.path` INSTEAD of the reference's `code/gym-lorenz`
(code/train.py:6-11 appends that path) and the reference scripts run unchanged:

    import gym_lorenz                      # registers lorenz_try-v0 / lorenz_pmsm-v0
    env_fn = lambda: gymnasium.make("lorenz_try-v0", add_noise=add_noise)
    env = DummyVecEnv([env_fn])            # code/train.py:98-100

The registered entry points are the GPU single-env facades of gym_lorenz_b200 (same constructor
kwargs, spaces, attributes); gymnasium adds the same TimeLimit (5000 / 2000 steps) on top.
For many envs at once use gym_lorenz_b200.vec_env.BatchedChaosVecEnv instead of DummyVecEnv.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from gym_lorenz_b200.envs import register_gymnasium  # noqa: E402

REGISTERED = register_gymnasium()   # False when gymnasium is not importable
