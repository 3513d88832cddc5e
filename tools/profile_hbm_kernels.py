#!/usr/bin/env python
"""A few launches of every HBM-bound kernel at the sizes their rooflines are quoted on, for one
`ncu --set full` capture (tools/r02_prof.sh): the single-step kernels of the parity kinds at 1 Mi envs
and the five device-side SB3 kernels of tu_rl_ops.cu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from gym_lorenz_b200 import rl_ops
from gym_lorenz_b200.core import ChaosBatch

kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ["memristive4_pair", "pmsm_free", "pmsm_classic", "pmsm_sync", "hr_sync"]
AMP = {"lorenz3": 0.05, "lorenz3_pair": 0.05, "memristive4_pair": 2.0, "pmsm_classic": 2.0, "pmsm_free": 0.0, "pmsm_single": 0.5}
n = 1 << 20
dev = torch.device("cuda:0")
for kind in kinds:
    if kind == "rl_ops":
        continue
    b = ChaosBatch(kind, n, seed=0)
    b.reset()
    a = (torch.rand((n, b.act_dim), device=dev) * 2 - 1) * AMP.get(kind, 1.0)
    for _ in range(2):
        b.step(a)
    torch.cuda.synchronize()
    b.close()
if "rl_ops" in kinds or len(sys.argv) <= 1:
    N, D, K = 1 << 20, 6, 4
    planes = torch.randn((D, N), device=dev)
    obs = planes.t()                                   # [N, D] view of SoA planes, as the env hands it over
    rms = rl_ops.RunningMeanStd((D,), dev)
    out = torch.empty((N, D), device=dev)
    stacked = torch.zeros((N, D * K), device=dev)
    term = torch.zeros_like(stacked)
    done = (torch.rand(N, device=dev) < 0.01).to(torch.uint8)
    lib = rl_ops.L.load()
    for _ in range(2):
        rms.update(obs)
        lib.cl_obs_normalize(rl_ops._stream(dev), rl_ops._p(obs), obs.stride(0), obs.stride(1), rl_ops._p(out), D, 1, N, D,
                             rl_ops._p(rms.mean), rl_ops._p(rms.var), 1e-8, 10.0)
        lib.cl_frame_stack_term(rl_ops._stream(dev), rl_ops._p(stacked), rl_ops._p(out), D, 1, rl_ops._p(done), rl_ops._p(out),
                                D, 1, rl_ops._p(term), N, D, K)
    T, NE = 128, 65536
    r, v, s = (torch.randn((T, NE), device=dev) for _ in range(3))
    for _ in range(2):
        rl_ops.gae(r, v, (s > 1).float(), v[0], (s[0] > 1).float(), 0.99, 0.95)
    e, u = torch.randn((2000, 3, 16384), device=dev, dtype=torch.float64), torch.randn((2000, 2, 16384), device=dev, dtype=torch.float64)
    for _ in range(2):
        rl_ops.eval_metrics(e, u, dt=0.001)
    torch.cuda.synchronize()
print("done")
