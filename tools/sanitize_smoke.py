#!/usr/bin/env python
"""Small, fast exercise of every kernel family for `compute-sanitizer --tool memcheck`:
ragged sizes (partial warps / blocks), static + dynamic rollout (bulk-copy staging), reset with
mask, host zero-copy path, RL ops."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_lorenz_b200.core import ChaosBatch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
from gym_lorenz_b200 import rl_ops

dev = torch.device("cuda:0")
for kind in ("lorenz3", "lorenz3_pair", "lorenz4_pair", "hr_sync", "pmsm_sync", "pmsm_classic", "pmsm_single",
             "lorenz_rk4", "lorenz_rk4_f32", "pmsm_rk4", "memristive4_pair", "pmsm_free"):
    for n in (1, 33, 1000):
        b = ChaosBatch(kind, n, seed=1, max_episode_steps=3, add_noise=kind in ("hr_sync", "pmsm_sync"))
        b.reset()
        a = torch.rand((n, b.act_dim), device=dev) * 0.1
        for _ in range(4):
            b.step(a)
        m = torch.zeros(n, dtype=torch.uint8, device=dev); m[::2] = 1
        b.reset(m)
        T = 19
        for dyn in ("0", "1"):
            os.environ["CHAOS_B200_DYN"] = dyn
            soa = torch.rand((T, b.act_dim, b.n_pad), device=dev) * 0.1
            b.rollout(T, soa[:, :, :n].permute(0, 2, 1))
            b.rollout(T, (torch.rand((T, n, b.act_dim), device=dev) * 0.1))
            b.rollout(T, None, want=("reward",))
        os.environ.pop("CHAOS_B200_DYN", None)
        b.stats()
        b.close()
for mode, slices in (("zerocopy", 1), ("dma", 1), ("pipelined", 3)):
    env = BatchedChaosVecEnv("hr_sync", 777, max_episode_steps=2)
    env.batch.set_host_mode(mode, slices)
    env.reset()
    for _ in range(5):
        obs, rew, dones, infos = env.step(np.zeros((777, 2), np.float32))   # episodes end: term_obs rows to the host slot
        _ = [d for d in infos if d]
    env.close()
r = torch.randn((16, 500), device=dev)
rl_ops.gae(r, r, (r > 1).float(), r[0], (r[1] > 0).float(), 0.99, 0.95)
rms = rl_ops.RunningMeanStd((6,), dev); rms.update(torch.randn((1000, 6), device=dev))
rl_ops.eval_metrics(torch.randn((50, 3, 77), device=dev, dtype=torch.float64), torch.randn((50, 2, 77), device=dev, dtype=torch.float64), 0.01)
torch.cuda.synchronize()
print("sanitize smoke ok")
