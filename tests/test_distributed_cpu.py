"""N>1 host logic on CPU: slab partitioning and the statistics all-reduce over gloo with
world_size 2 (the same code path bench.py / ChaosBatch use with NCCL on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_lorenz_b200 import distributed as D


def test_partition_covers_all_envs_without_overlap():
    for total, world in ((65536, 8), (1000, 3), (7, 8), (8 * 1048576, 8)):
        seen = 0
        for r in range(world):
            s = D.partition(total, world, r)
            assert s.env_id_base == seen
            seen += s.num_envs
        assert seen == total
    s = D.weak_slab(1048576, 8, 5)
    assert s.env_id_base == 5 * 1048576 and s.total_envs == 8 * 1048576


def test_summarize():
    out = D.summarize([4.0, 10.0, 30.0, 40.0, 0, 1, 3, 0])
    assert out["ep_rew_mean"] == 2.5 and out["ep_len_mean"] == 10.0
    assert np.isclose(out["ep_rew_std"], np.sqrt(30 / 4 - 2.5 ** 2))
    assert np.isnan(D.summarize([0.0] * 8)["ep_rew_mean"])


def test_rank_affinity_plan():
    """Each rank gets its own contiguous share of the cores of its GPU's NUMA node."""
    node_cpus = {0: list(range(0, 48)) + list(range(96, 144)), 1: list(range(48, 96)) + list(range(144, 192))}
    allowed = list(range(192))
    gpu_nodes = [0, 0, 0, 0, 1, 1, 1, 1]
    plans = [D.plan_affinity(allowed, gpu_nodes, node_cpus, r) for r in range(8)]
    assert all(len(p) == 24 for p in plans)
    assert sorted(c for p in plans[:4] for c in p) == sorted(node_cpus[0])
    assert sorted(c for p in plans[4:] for c in p) == sorted(node_cpus[1])
    assert len({c for p in plans for c in p}) == 192                  # disjoint
    # restricted mask (container cpuset): only allowed cores are handed out
    allowed = list(range(8, 40))
    plans = [D.plan_affinity(allowed, [0, 0], node_cpus, r) for r in range(2)]
    assert plans[0] == list(range(8, 24)) and plans[1] == list(range(24, 40))
    # unknown topology: even split of the allowed set over all ranks
    plans = [D.plan_affinity(list(range(10)), [-1, -1, -1], {}, r) for r in range(3)]
    assert plans == [[0, 1, 2, 3], [4, 5, 6], [7, 8, 9]]
    # a node with fewer allowed cores than ranks: fall back to sharing everything, never empty
    plans = [D.plan_affinity([0, 1, 50, 51], [0, 0, 0, 1], {0: [0], 1: [50, 51]}, r) for r in range(4)]
    assert all(len(p) >= 1 for p in plans)
    assert D.plan_affinity([3, 4], [0], {0: [3, 4]}, 0) == [3, 4]      # single rank: untouched
    assert D._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = D.init_process_group("gloo")
    slab = D.weak_slab(1000, w, r)
    # each rank's local statistics (as ChaosBatch.stats_tensor() would return them)
    local = torch.tensor([slab.num_envs, 2.0 * (r + 1), 4.0 * (r + 1), 10.0 * slab.num_envs, r, 0, slab.num_envs, 0],
                         dtype=torch.float64)
    D.allreduce_stats(local)
    q.put((r, slab.env_id_base, local.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_allreduce_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [0, 1000]
    expect = [2000.0, 6.0, 12.0, 20000.0, 1.0, 0.0, 2000.0, 0.0]
    assert res[0][2] == expect and res[1][2] == expect


def test_allreduce_is_noop_single_process():
    t = torch.arange(8, dtype=torch.float64)
    assert D.allreduce_stats(t) is None and t.tolist() == list(range(8))
    with pytest.raises(ValueError):
        D.allreduce_stats(torch.zeros(3, dtype=torch.float64))
