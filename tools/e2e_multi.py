#!/usr/bin/env python
"""Host-path (SB3 VecEnv contract, caller-owned action array every step) step time at N ranks, one rank per
GPU, for several placements / data-movement modes -- what limits the e2e number at 8 GPUs.
  torchrun --nproc-per-node N tools/e2e_multi.py        (prints one JSON line per variant, max over ranks)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from gym_lorenz_b200 import distributed as D
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv

rank, world, local = D.init_process_group()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
all_cores = sorted(os.sched_getaffinity(0))
nodes, cpus = D.gpu_numa_topology(world)
mine = D.plan_affinity(all_cores, nodes, cpus, local)
if rank == 0:
    print(json.dumps({"world": world, "logical_cpus": len(all_cores), "gpu_numa_nodes": nodes,
                      "node_cpu_counts": {k: len(v) for k, v in cpus.items()}, "rank0_cores": mine}), flush=True)
N = 65536
kind = sys.argv[1] if len(sys.argv) > 1 else "lorenz_rk4"
variants = [("pin", "streamed", 2), ("pin", "streamed", 1), ("pin", "zerocopy", 1), ("pin", "dma", 1),
            ("nopin", "streamed", 2), ("nopin", "streamed", 1), ("nopin", "zerocopy", 1), ("pin", "pinned-input", 1)]
for place, mode, threads in variants:
    os.sched_setaffinity(0, mine if place == "pin" else all_cores)
    os.environ["CHAOS_B200_COPY_THREADS"] = str(threads)
    env = BatchedChaosVecEnv(kind, N, device=dev, seed=0, env_id_base=rank * N, max_episode_steps=1000)
    b = env.batch
    env.reset()
    if mode != "pinned-input":
        b.set_host_mode(mode, 32 if mode == "streamed" else 1)
    rng = np.random.default_rng(rank)
    acts = [rng.uniform(-1, 1, (N, b.act_dim)).astype(np.float32) for _ in range(8)]
    pin = b.host_action_buffer(); pin[:] = acts[0]
    step = (lambda k: env.step(pin)) if mode == "pinned-input" else (lambda k: env.step(acts[k % 8]))
    for k in range(100):
        step(k)
    dist.barrier(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for k in range(400):
        step(k)
    torch.cuda.synchronize(dev)
    us = (time.perf_counter() - t0) / 400 * 1e6
    t = torch.tensor([us], dtype=torch.float64, device=dev)
    tmin = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"kind": kind, "envs_per_gpu": N, "world": world, "placement": place, "mode": mode, "copy_threads": threads,
                          "us_per_step_max_over_ranks": round(float(t.item()), 2), "us_per_step_min_over_ranks": round(float(tmin.item()), 2),
                          "env_steps_per_s_all_ranks": N * world / (float(t.item()) * 1e-6)}), flush=True)
    env.close()
    dist.barrier()
dist.destroy_process_group()
