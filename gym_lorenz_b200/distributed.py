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
from typing import Dict, List, Optional, Sequence, Tuple

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


# ---- CPU placement of the ranks --------------------------------------------------------------
# The SB3-facing host path (step_async / step_wait with host buffers) is driven by one Python thread
# per rank and moves ~2.7 MB per step through pinned host memory.  With 8 ranks on one box and no
# placement, the ranks' threads migrate over all cores of both sockets and their pinned buffers end up
# on whichever NUMA node first touched them: measured e2e scaling efficiency 0.53 at 8 GPUs (round 1).
# Each rank therefore pins itself to its own share of the cores of the NUMA node its GPU hangs off,
# BEFORE it allocates pinned memory (first touch then places the buffers on that node) -- when sysfs
# reports that node.  Where it does not (numa_node = -1), ranks are left to the scheduler.

def _parse_cpulist(text: str) -> List[int]:
    out: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            out.extend(range(int(lo), int(hi) + 1))
        else:
            out.append(int(part))
    return out


def plan_affinity(allowed: Sequence[int], gpu_nodes: Sequence[int], node_cpus: Dict[int, Sequence[int]],
                  local_rank: int) -> List[int]:
    """Cores for `local_rank` (one rank per GPU, rank r drives GPU r).  `gpu_nodes[r]` is the NUMA node of
    GPU r (-1 = unknown), `node_cpus[n]` the logical CPUs of node n, `allowed` this process's current
    affinity mask.  Ranks whose GPUs share a node split that node's allowed cores contiguously and
    evenly; unknown topology falls back to an even split of `allowed` over all ranks.  Never returns
    an empty set."""
    allowed = sorted(set(int(c) for c in allowed))
    world = len(gpu_nodes)
    if world <= 1 or not allowed:
        return allowed
    node = gpu_nodes[local_rank]
    local = sorted(set(node_cpus.get(node, ())) & set(allowed)) if node >= 0 else []
    peers = [r for r in range(world) if gpu_nodes[r] == node]
    if not local or len(local) < len(peers):
        local, peers = allowed, list(range(world))      # no usable topology: share everything evenly
    k = peers.index(local_rank)
    q, rem = divmod(len(local), len(peers))
    lo = k * q + min(k, rem)
    hi = lo + q + (1 if k < rem else 0)
    return local[lo:hi] if hi > lo else [local[k % len(local)]]


def gpu_numa_topology(world: int) -> Tuple[List[int], Dict[int, List[int]]]:
    """(NUMA node of each of the first `world` GPUs, cpu list per node) from sysfs; -1 / {} where unknown."""
    nodes: List[int] = []
    cpus: Dict[int, List[int]] = {}
    for r in range(world):
        node = -1
        try:
            pr = torch.cuda.get_device_properties(r)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
                node = int(f.read().strip())
        except Exception:  # noqa: BLE001
            node = -1
        nodes.append(node)
        if node >= 0 and node not in cpus:
            try:
                with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                    cpus[node] = _parse_cpulist(f.read())
            except Exception:  # noqa: BLE001
                pass
    return nodes, cpus


def pin_rank_to_cores(local_rank: int, local_world: int) -> List[int]:
    """Pin the calling process (all its current threads inherit on creation) to this rank's cores.
    Call it before the first pinned allocation.  Returns the cores chosen (the old mask if world is 1)."""
    allowed = sorted(os.sched_getaffinity(0))
    if local_world <= 1:
        return allowed
    nodes, cpus = gpu_numa_topology(local_world)
    if local_rank >= len(nodes) or nodes[local_rank] < 0 or nodes[local_rank] not in cpus:
        # no topology (e.g. a virtualised box reports numa_node = -1): an even split buys nothing -- measured at
        # 8 GPUs / 32 logical CPUs: 185 us per step pinned vs 178 us unpinned (profiles/r02_e2e_multi_8gpu.jsonl)
        return allowed
    mine = plan_affinity(allowed, nodes, cpus, local_rank)
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return allowed
    # intra-op thread pools sized for the whole box would oversubscribe the rank's share
    torch.set_num_threads(max(1, min(torch.get_num_threads(), len(mine))))
    return mine


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
