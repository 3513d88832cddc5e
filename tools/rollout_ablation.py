#!/usr/bin/env python
"""Where the rollout kernel's time goes beyond the RK4 substeps: same workload as bench.py with
output streams / action source switched off one at a time."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from gym_lorenz_b200.core import ChaosBatch, measure_fma_peak

N, T, S = 65536, 256, 16
peak = measure_fma_peak(0, 8, 0.2)
dev = torch.device("cuda:0")
def run(want, actions_mode, substeps=S, reps=40):
    b = ChaosBatch("lorenz_rk4", N, seed=0, substeps=substeps, max_episode_steps=1000)
    b.reset()
    g = torch.Generator(device=dev).manual_seed(0)
    soa = torch.rand((T, 3, b.n_pad), generator=g, device=dev) * 2 - 1
    acts = {"soa": soa[:, :, :N].permute(0, 2, 1), "aos": soa[:, :, :N].permute(0, 2, 1).contiguous(), "philox": None}[actions_mode]
    out = b.rollout(T, acts, want=want)
    for _ in range(3): b.rollout(T, acts, out=out, want=want)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): b.rollout(T, acts, out=out, want=want)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    b.close()
    return ms
base = None
for name, want, am, sub in [("full (obs+reward+done, SoA actions via TMA)", ("obs", "reward", "done"), "soa", S),
                            ("no obs stream", ("reward", "done"), "soa", S),
                            ("no outputs at all", (), "soa", S),
                            ("no outputs, Philox actions", (), "philox", S),
                            ("full, AoS actions (LDG path)", ("obs", "reward", "done"), "aos", S),
                            ("full, S=32", ("obs", "reward", "done"), "soa", 32),
                            ("full, S=8", ("obs", "reward", "done"), "soa", 8)]:
    ms = run(want, am, sub)
    tf = N * T * sub * 87 / (ms * 1e-3) * 1e-12
    print(json.dumps({"variant": name, "ms_per_launch": round(ms, 4), "tflops": round(tf, 2), "frac_of_dfma_peak": round(tf / peak, 4)}), flush=True)
