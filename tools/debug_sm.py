import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from gym_lorenz_b200.core import ChaosBatch
n, T = int(sys.argv[1]), int(sys.argv[2])
b = ChaosBatch("lorenz_rk4", n, seed=1, max_episode_steps=11)
b.reset()
acts = (torch.rand((T, 3, b.n_pad), device=b.device) * 2 - 1)
try:
    out = b.rollout(T, acts[:, :, :n].permute(0, 2, 1))
    torch.cuda.synchronize()
    print("OK", n, T, "sm", b.sm_launch_count, "dyn", b.dyn_launch_count, float(out["reward"][:, :n].sum()))
except Exception as e:
    print("FAIL", n, T, str(e)[:100])
