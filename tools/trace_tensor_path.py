#!/usr/bin/env python
"""torch.profiler (CUPTI) trace summary of the tensor / DLPack path: 64 policy -> step_tensor iterations
at 4,096 HR envs; prints per-activity counts (kernels by name, memcpys by direction) as JSON."""
import collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv

n = 4096
env = BatchedChaosVecEnv("hr_sync", n, seed=3)
net = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.Tanh(), torch.nn.Linear(32, 2), torch.nn.Tanh()).to("cuda:0")
obs = env.reset_tensor()
with torch.no_grad():
    for _ in range(4):
        obs, rew, done = env.step_tensor(net(obs))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(64):
            obs, rew, done = env.step_tensor(net(obs))
        torch.cuda.synchronize()
cnt = collections.Counter()
for e in prof.events():
    if str(e.device_type).endswith("CUDA"):
        cnt[e.name[:90]] += 1
memcpy = {k: v for k, v in cnt.items() if "memcpy" in k.lower() or "memset" in k.lower()}
print(json.dumps({"envs": n, "iterations": 64, "cuda_activity_counts": dict(cnt.most_common(12)),
                  "memcpy_memset_records": memcpy,
                  "host_device_copies_in_loop": sum(v for k, v in memcpy.items() if "htod" in k.lower() or "dtoh" in k.lower())}))
