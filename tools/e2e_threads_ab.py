#!/usr/bin/env python
"""A/B of the staging lanes (copy threads 1..4) and slice counts inside ONE process (A B C D A B C D ...:
successive processes land on different cores and differ by +-12 % on their own): streamed mode, caller-owned
array every step.  argv: kind [threads:slices,...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
kind = sys.argv[1] if len(sys.argv) > 1 else "lorenz_rk4"
N = 65536
rng = np.random.default_rng(0)
variants = [(1, 32), (2, 32), (3, 33), (4, 32), (4, 64), (4, 16)]
if len(sys.argv) > 2:
    variants = [tuple(int(x) for x in v.split(":")) for v in sys.argv[2].split(",")]
for rep in range(3):
    for th, sl in variants:
        os.environ["CHAOS_B200_COPY_THREADS"] = str(abs(th))
        os.environ["CHAOS_B200_RELAY"] = "kernel" if th < 0 else "block"      # negative thread count: k_relay as its own launch
        env = BatchedChaosVecEnv(kind, N)
        env.batch.set_host_mode("streamed", sl)
        env.reset()
        acts = [rng.uniform(-1, 1, (N, env.batch.act_dim)).astype(np.float32) for _ in range(8)]
        for k in range(60):
            env.step(acts[k % 8])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(400):
            env.step(acts[k % 8])
        torch.cuda.synchronize()
        row = {"kind": kind, "envs": N, "rep": rep, "copy_threads": abs(th), "relay": "kernel" if th < 0 else "block0", "slices": sl,
               "us_per_step": round((time.perf_counter() - t0) / 400 * 1e6, 2)}
        if (th, sl) == variants[-1]:        # the floor: actions already in the pinned staging buffer
            env.batch.host_action_buffer()[:] = acts[0]
            for k in range(60):
                env.batch.step_host_async(None); env.batch.step_host_wait()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for k in range(400):
                env.batch.step_host_async(None); env.batch.step_host_wait()
            torch.cuda.synchronize()
            row["us_per_step_pinned_actions"] = round((time.perf_counter() - t0) / 400 * 1e6, 2)
        print(json.dumps(row), flush=True)
        env.close()
