#!/usr/bin/env python
"""Why is the zero-copy host step slow for hr_sync at 65,536 envs?  async / wait split, n_done."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
for kind in ("hr_sync", "lorenz_rk4", "pmsm_sync"):
    for N in (4096, 65536):
        env = BatchedChaosVecEnv(kind, N); env.reset(); b = env.batch
        pin = b.host_action_buffer(); pin[:] = np.random.default_rng(0).uniform(-1, 1, pin.shape).astype(np.float32)
        for mode, s in (("dma", 1), ("zerocopy", 1), ("pipelined", 1)):
            b.set_host_mode(mode, s)
            ta = tw = 0.0; nd = 0
            for k in range(330):
                t0 = time.perf_counter(); b.step_host_async(None); t1 = time.perf_counter()
                r = b.step_host_wait(); t2 = time.perf_counter()
                if k >= 30: ta += t1 - t0; tw += t2 - t1; nd += int(r[-1] > 0)
            print(f"{kind:12s} N={N:6d} {mode:10s} async {ta/300*1e6:6.1f} us  wait {tw/300*1e6:6.1f} us  steps with done: {nd}/300", flush=True)
        env.close()
