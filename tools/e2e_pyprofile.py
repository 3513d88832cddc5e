#!/usr/bin/env python
"""Split of one host-buffer step into its calls (perf_counter_ns around each), 65,536 Lorenz envs."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np, torch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = BatchedChaosVecEnv("lorenz_rk4", N); env.reset()
b = env.batch
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (N, 3)).astype(np.float32) for _ in range(8)]
ns = time.perf_counter_ns
for label, use_pin in (("ndarray", False), ("pinned", True)):
    pin = b.host_action_buffer(); pin[:] = acts[0]
    for k in range(50): env.step(pin if use_pin else acts[k % 8])
    ta = tw = tt = tl = 0
    lib, ctx, st = b.lib, b.ctx, b._host_stream()
    for k in range(400):
        a = pin if use_pin else acts[k % 8]
        t0 = ns(); env.step_async(a); t1 = ns(); out = env.step_wait(); t2 = ns()
        ta += t1 - t0; tw += t2 - t1
    # raw library calls only (no Python wrappers beyond ctypes)
    ap = None
    for k in range(400):
        a = acts[k % 8]
        p = None if use_pin else a.ctypes.data
        t0 = ns(); lib.cl_step_host_async(ctx, st, b._bufs_ref, p); t1 = ns()
        lib.cl_step_host_wait_view(ctx, st, b._host_view_ref); t2 = ns()
        tl += t1 - t0; tt += t2 - t1
    print(json.dumps({"actions": label, "envs": N, "us_step_async_python": round(ta / 400e3, 2), "us_step_wait_python": round(tw / 400e3, 2),
                      "us_cl_step_host_async_raw": round(tl / 400e3, 2), "us_cl_step_host_wait_raw": round(tt / 400e3, 2)}), flush=True)
env.close()
