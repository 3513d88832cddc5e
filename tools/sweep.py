#!/usr/bin/env python
"""Throughput of every env kind on one B200: single-step API (`cl_step`, one
launch per control interval, device-resident actions) and fused rollout (`cl_rollout`), at
65,536 and 1,048,576 envs.  Prints one JSON line per measurement with the roofline that bounds it
(HBM for the parity kinds, FP64/FP32 FMA for RK4 x S).  Not part of the bench contract; output is
kept under profiles/."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from gym_lorenz_b200.core import ChaosBatch, measure_fma_peak  # noqa: E402

# kind: (kwargs, action amplitude, algorithmic bytes per env-step for the single-step API,
#        algorithmic flop per env-step)
KINDS = {
    "lorenz3": ({}, 0.05, 32 + 12 + 32 + 24 + 8 + 1 + 12, 31),            # state+t rw, act, obs f32, reward, done, ep counters
    "lorenz3_pair": ({}, 0.05, 80 + 12 + 80 + 24 + 8 + 1 + 12, 37),
    "lorenz4_pair": ({}, 1.0, 72 + 12 + 72 + 32 + 8 + 1 + 12, 81),
    "hr_sync": ({}, 1.0, 72 + 8 + 72 + 24 + 8 + 1 + 12, 250),
    "pmsm_sync": ({"alpha": 0.5}, 1.0, 40 + 8 + 40 + 24 + 4 + 1 + 12, 120),
    "pmsm_classic": ({}, 2.0, 56 + 8 + 56 + 24 + 8 + 1 + 12, 60),
    "pmsm_single": ({}, 0.5, 32 + 8 + 32 + 24 + 8 + 1 + 12, 31),
    "lorenz_rk4": ({"substeps": 16}, 1.0, 48 + 12 + 48 + 24 + 8 + 1 + 12, 16 * 87),
    "lorenz_rk4_f32": ({"substeps": 16}, 1.0, 24 + 12 + 24 + 24 + 4 + 1 + 12, 16 * 87),
    "pmsm_rk4": ({"substeps": 4}, 1.0, 64 + 8 + 64 + 24 + 8 + 1 + 12, 4 * 2 * 91),
    "memristive4_pair": ({}, 2.0, 72 + 12 + 72 + 32 + 8 + 1 + 12, 95),
    "pmsm_free": ({}, 0.0, 32 + 8 + 32 + 24 + 8 + 1 + 12, 34),
}


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="65536,1048576")
    ap.add_argument("--kinds", default=",".join(KINDS))
    ap.add_argument("--T", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    fp64 = measure_fma_peak(0, 8, 0.3)
    fp32 = measure_fma_peak(0, 4, 0.3)
    print(json.dumps({"peaks": {"hbm_gbs": hbm, "fp64_tflops_measured": fp64, "fp32_tflops_measured": fp32}}), flush=True)
    for kind in args.kinds.split(","):
        kw, amp, bytes_step, flop = KINDS[kind]
        for n in (int(x) for x in args.sizes.split(",")):
            b = ChaosBatch(kind, n, seed=0, autoreset=True, **kw)
            b.reset()
            T = args.T
            g = torch.Generator(device=dev).manual_seed(0)
            soa = (torch.rand((T, b.act_dim, b.n_pad), generator=g, device=dev) * 2 - 1) * amp
            acts = soa[:, :, :n].permute(0, 2, 1)
            a0 = acts[0]
            # single steps are timed as a CUDA-graph replay of 20 launches (graph mode keeps the Philox
            # step index on the device), so the number is kernel time, not Python/ctypes call overhead
            b.set_graph_mode(True)
            for _ in range(3):
                b.step(a0)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    b.step(a0)
            ms_step = timed(g.replay, 20 if n <= 65536 else 5) / 20
            ms_step_eager = timed(lambda: b.step(a0), 100 if n <= 65536 else 30)
            out = b.rollout(T, acts)
            ms_roll = timed(lambda: b.rollout(T, acts, out=out), 20 if n <= 65536 else 5)
            peak_tf = fp32 if b.real == torch.float32 and kind != "pmsm_sync" else fp64
            for mode, ms, steps, byts in (("step", ms_step, 1, bytes_step),
                                          ("rollout", ms_roll, T, b.act_dim * 4 + b.obs_dim * 4 + b.layout.real_bytes + 1)):
                rate = n * steps / (ms * 1e-3)
                print(json.dumps({
                    "kind": kind, "n": n, "mode": mode, "T": steps, "ms_per_launch": round(ms, 5),
                    "env_steps_per_s": rate, "gbs": rate * byts * 1e-9, "hbm_frac": rate * byts * 1e-9 / hbm,
                    "tflops": rate * flop * 1e-12, "fma_frac": rate * flop * 1e-12 / peak_tf,
                    "dyn": b.dyn_launch_count > 0, "block": b.block_size,
                    "eager_python_ms_per_step": round(ms_step_eager, 5) if mode == "step" else None}), flush=True)
            b.close()
            del b, soa, acts, out
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
