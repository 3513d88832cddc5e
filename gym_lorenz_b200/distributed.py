"""Multi-GPU plumbing: one process per GPU, one independent slab of envs per rank.

Envs never interact (each `step` touches only its own state, e.g. dynamic.py:69-76), so the
data path has NO collective.  The only exchange is an optional all-reduce (sum) of the
8-double episode-statistics vector -- 64 bytes, latency-bound -- via torch.distributed
(NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).  Philox subsequences are keyed by
the GLOBAL env index (env_id_base + local index), so a given env sees the same random numbers
whichever rank owns it: results are independent of the GPU count.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import NSTATS, STAT_NAMES


@dataclass(frozen=True)
class Slab:
    rank: int
    world_size: int
    num_envs: int       # envs owned by this rank
    env_id_base: int    # global index of local env 0
    total_envs: int


def partition(total_envs: int, world_size: int, rank: int) -> Slab:
    """Contiguous slabs; the first `total % world` ranks take one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    q, r = divmod(int(total_envs), int(world_size))
    n = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return Slab(rank, world_size, n, base, int(total_envs))


def weak_slab(envs_per_rank: int, world_size: int, rank: int) -> Slab:
    return Slab(rank, world_size, int(envs_per_rank), rank * int(envs_per_rank), world_size * int(envs_per_rank))


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment (no-op for world 1)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def allreduce_stats(stats: torch.Tensor, async_op: bool = False):
    """Sum the CL_NSTATS vector over ranks in place (no-op when not distributed)."""
    if stats.numel() != NSTATS:
        raise ValueError(f"expected {NSTATS} statistics, got {stats.numel()}")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(stats, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


def summarize(stats) -> Dict[str, float]:
    """episode count / mean / std of return / mean length from the raw sums."""
    v = stats.tolist() if hasattr(stats, "tolist") else list(stats)
    d = dict(zip(STAT_NAMES, v))
    n = d["episodes"]
    out = dict(d)
    if n > 0:
        mean = d["return_sum"] / n
        var = max(d["return_sq_sum"] / n - mean * mean, 0.0)
        out.update(ep_rew_mean=mean, ep_rew_std=math.sqrt(var), ep_len_mean=d["length_sum"] / n)
    else:
        out.update(ep_rew_mean=float("nan"), ep_rew_std=float("nan"), ep_len_mean=float("nan"))
    return out
