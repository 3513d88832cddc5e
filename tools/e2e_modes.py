#!/usr/bin/env python
"""Host-buffer VecEnv step time vs data-movement mode (cl_host_set_mode): us per step."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv

kind = sys.argv[1] if len(sys.argv) > 1 else "lorenz_rk4"
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096, 16384, 65536, 262144]
modes = [("dma", 1), ("zerocopy", 1), ("pipelined", 2), ("pipelined", 4), ("streamed", 1), ("streamed", 2),
         ("streamed", 4), ("streamed", 8), ("streamed", 16)]
if len(sys.argv) > 3:
    modes = [(m.split(":")[0], int(m.split(":")[1])) for m in sys.argv[3].split(",")]
for N in sizes:
    env = BatchedChaosVecEnv(kind, N)
    env.reset()
    b = env.batch
    rng = np.random.default_rng(0)
    # 4 rotating arrays stay warm in the CPU caches (a policy that has just written its output); E2E_ARRAYS=16
    # makes the source cold (L3 / DRAM)
    n_arr = int(os.environ.get("E2E_ARRAYS", "4"))
    acts = [rng.uniform(-1, 1, (N, b.act_dim)).astype(np.float32) for _ in range(n_arr)]
    pin = b.host_action_buffer(); pin[:] = acts[0]

    def bench(fn, n=300):
        for _ in range(30): fn(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(n): fn(k)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

    def pinned(k):
        b.step_host_async(None); b.step_host_wait()
    for mode, s in modes:
        b.set_host_mode(mode, s)
        row = {"kind": kind, "envs": N, "mode": mode, "slices": s,
               "us_pinned_actions": round(bench(pinned), 2),
               "us_ndarray_actions": round(bench(lambda k: env.step(acts[k % n_arr])), 2), "arrays": n_arr,
               "copy_threads": int(os.environ.get("CHAOS_B200_COPY_THREADS", "1"))}
        print(json.dumps(row), flush=True)
    env.close()
