#!/usr/bin/env python
"""Zero-copy vs streamed host mode at small batches, A-B-A-B on ONE env in one process (caller-owned array each step)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
for kind, N in (("lorenz_rk4", 4096), ("lorenz_rk4", 8192), ("lorenz_rk4", 16384), ("hr_sync", 8192), ("hr_sync", 16384), ("pmsm_sync", 16384)):
    env = BatchedChaosVecEnv(kind, N); env.reset()
    rng = np.random.default_rng(0)
    acts = [rng.uniform(-0.3, 0.3, (N, env.batch.act_dim)).astype(np.float32) for _ in range(8)]
    res = {}
    for rep in range(3):
        for mode, sl in (("zerocopy", 1), ("streamed", max(1, N // 2048))):
            env.batch.set_host_mode(mode, sl)
            for k in range(40): env.step(acts[k % 8])
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for k in range(300): env.step(acts[k % 8])
            torch.cuda.synchronize()
            res.setdefault(mode, []).append(round((time.perf_counter() - t0) / 300 * 1e6, 2))
    print(json.dumps({"kind": kind, "envs": N, "action_kb": N * env.batch.act_dim * 4 // 1024, **res}), flush=True)
    env.close()
