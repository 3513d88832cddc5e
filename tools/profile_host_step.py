#!/usr/bin/env python
"""A few host-buffer steps in zero-copy mode at 65,536 Lorenz envs (for an ncu capture of the e2e step kernel:
`ncu --set full -k regex:k_step -s 12 -c 1 python tools/profile_host_step.py`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
env = BatchedChaosVecEnv("lorenz_rk4", 65536)
env.batch.set_host_mode("zerocopy", 1)
env.reset()
rng = np.random.default_rng(0)
a = rng.uniform(-1, 1, (65536, 3)).astype(np.float32)
for _ in range(20):
    env.step(a)
env.close()
