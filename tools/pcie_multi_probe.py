#!/usr/bin/env python
"""Is the host path's slowdown at 8 ranks ours or the box's?  Pure copy-engine traffic of one host-buffer step
(1.9 MB device->host + 0.79 MB host->device, pinned, two streams) per rank: first rank 0 alone, then all ranks
at once.   torchrun --nproc-per-node N tools/pcie_multi_probe.py   (prints one JSON line)"""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
D2H, H2D = 65536 * 29, 65536 * 12
d_out, h_out = torch.empty(D2H, dtype=torch.uint8, device=dev), torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_in, h_in = torch.empty(H2D, dtype=torch.uint8, device=dev), torch.empty(H2D, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def one():
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)
    s1.synchronize(); s2.synchronize()


def bench(n=2000):
    for _ in range(100):
        one()
    t0 = time.perf_counter()
    for _ in range(n):
        one()
    return (time.perf_counter() - t0) / n * 1e6


def barrier():
    if world > 1:
        dist.barrier()


alone = None
barrier()
if rank == 0:
    alone = bench()
barrier()
together = bench()
barrier()
t = torch.tensor([together, -together], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"ranks": world, "bytes_d2h": D2H, "bytes_h2d": H2D, "us_rank0_alone": round(alone, 2),
                      "us_all_ranks_max": round(float(t[0]), 2), "us_all_ranks_min": round(-float(t[1]), 2)}), flush=True)
if world > 1:
    dist.destroy_process_group()
