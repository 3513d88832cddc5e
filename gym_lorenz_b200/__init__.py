"""gym_lorenz_b200 -- B200-native batched chaos-control environments.

The one hot path of erererq/gym-lorenz (env step: ODE integration -> reward -> termination
-> auto-reset) as hand-written sm_100a CUDA kernels behind the reference's own gym /
gymnasium / Stable-Baselines3 VecEnv surface.  See DESIGN.md and INTEGRATION.md.
"""
from ._lib import ChaosLibError, KIND_NAMES  # noqa: F401

__all__ = ["ChaosBatch", "BatchedChaosVecEnv", "ChaosLibError", "KIND_NAMES", "measure_fma_peak"]
__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not require torch/CUDA
    if name in ("ChaosBatch", "measure_fma_peak"):
        from . import core
        return getattr(core, name)
    if name == "BatchedChaosVecEnv":
        from .vec_env import BatchedChaosVecEnv
        return BatchedChaosVecEnv
    raise AttributeError(name)
