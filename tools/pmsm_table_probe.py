#!/usr/bin/env python
"""Does the Adam bias-correction table gather cost pmsm_sync's single step anything?  Same step with the per-env
Adam counter inside the tables (fresh envs) and beyond them (both corrections are exactly 1.0f, no lookups).
Measured at 1 Mi envs on B200: 28.6 / 28.0 / 28.0 us with the counter at 0 / 200 / 20,000 -- the gather is not the limiter."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from gym_lorenz_b200.core import ChaosBatch
n = 1 << 20
for adam0 in (0, 200, 20000):
    b = ChaosBatch("pmsm_sync", n, seed=0, autoreset=True, alpha=0.5); b.reset()
    b.aux_int[0, :] = adam0
    a0 = (torch.rand((n, 2), device=b.device) * 2 - 1)
    b.set_graph_mode(True)
    for _ in range(3): b.step(a0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): b.step(a0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"adam0": adam0, "us_per_step": round(e0.elapsed_time(e1) / 100 * 1e3, 2)}), flush=True)
    b.set_graph_mode(False); b.close()
