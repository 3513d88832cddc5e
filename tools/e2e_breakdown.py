#!/usr/bin/env python
"""Where the time of one host-buffer VecEnv step goes (N = 65,536 Lorenz RK4x16)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kind = sys.argv[2] if len(sys.argv) > 2 else "lorenz_rk4"
env = BatchedChaosVecEnv(kind, N)
env.reset()
b = env.batch
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (N, b.act_dim)).astype(np.float32) for _ in range(8)]

def bench(fn, n=300):
    for _ in range(20): fn(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n): fn(k)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

print("full env.step(np array)            %.1f us" % bench(lambda k: env.step(acts[k % 8])))
pin = b.host_action_buffer(); pin[:] = acts[0]
def pre(k):
    b.step_host_async(None); return b.step_host_wait()
print("actions already in pinned staging  %.1f us" % bench(pre))
def raw(k):
    b.step_host_async(None); b.step_host_wait()
print("memcpy user->pinned only           %.1f us" % bench(lambda k: np.copyto(pin, acts[k % 8])))
d_act = torch.as_tensor(acts[0], device=b.device)
print("device-resident cl_step only       %.1f us" % bench(lambda k: b.step(d_act)))
hp = torch.empty((N, b.act_dim), dtype=torch.float32).pin_memory(); dd = torch.empty_like(hp, device=b.device)
def h2d(k):
    dd.copy_(hp, non_blocking=True); torch.cuda.synchronize()
print("H2D %.0f KB pinned + sync           %.1f us" % (hp.numel() * 4 / 1e3, bench(h2d)))
out_bytes = N * (b.obs_dim * 4 + 4 + 1)
ho = torch.empty(out_bytes, dtype=torch.uint8).pin_memory(); do = torch.empty(out_bytes, dtype=torch.uint8, device=b.device)
def d2h(k):
    ho.copy_(do, non_blocking=True); torch.cuda.synchronize()
print("D2H %.0f KB pinned + sync          %.1f us" % (out_bytes / 1e3, bench(d2h)))
print("bare torch.cuda.synchronize        %.1f us" % bench(lambda k: torch.cuda.synchronize()))
env.close()
