#!/usr/bin/env python
"""A/B of the second staging thread inside ONE process (A B A B ...: successive processes land on different
cores and differ by +-12 % on their own): streamed mode, 32 slices, caller-owned array every step."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
kind = sys.argv[1] if len(sys.argv) > 1 else "lorenz_rk4"
N = 65536
rng = np.random.default_rng(0)
for rep in range(4):
    for th in (1, 2):
        os.environ["CHAOS_B200_COPY_THREADS"] = str(th)
        env = BatchedChaosVecEnv(kind, N)
        env.reset()
        acts = [rng.uniform(-1, 1, (N, env.batch.act_dim)).astype(np.float32) for _ in range(8)]
        for k in range(60):
            env.step(acts[k % 8])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(400):
            env.step(acts[k % 8])
        torch.cuda.synchronize()
        print(json.dumps({"kind": kind, "envs": N, "rep": rep, "copy_threads": th,
                          "us_per_step": round((time.perf_counter() - t0) / 400 * 1e6, 2)}), flush=True)
        env.close()
