#!/usr/bin/env python
"""A few single-step launches of the reference's own schemes (dynamic.py Lorenz, HRSyncEnv) at
1 Mi envs, for an ncu capture of the HBM-bound kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from gym_lorenz_b200.core import ChaosBatch
n = 1 << 20
for kind, amp in (("lorenz3", 0.05), ("hr_sync", 1.0), ("pmsm_sync", 1.0)):
    b = ChaosBatch(kind, n, seed=0)
    b.reset()
    a = (torch.rand((n, b.act_dim), device=b.device) * 2 - 1) * amp
    for _ in range(6):
        b.step(a)
    torch.cuda.synchronize()
    b.close()
print("done")
