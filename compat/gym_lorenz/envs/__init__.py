"""Same names as the reference's gym_lorenz/envs/__init__.py:2-3, GPU-backed."""
from gym_lorenz_b200.envs import HRSyncEnv, PMSM_Sync_Env  # noqa: F401
from gym_lorenz_b200.envs import lorenzEnv_transient  # noqa: F401  (dynamic.py; commented out upstream)
