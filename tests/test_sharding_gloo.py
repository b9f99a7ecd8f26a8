"""N > 1 host logic on CPU: a fixed global batch is dealt out over the ranks (round robin, strong scaling) with no data-path
collective; per-rank counts and times are combined exactly as bench.py does (max of times, sum of counts).  world_size = 2, gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conflict_rez_b200.control.batch_planner import random_init_offsets, shard_instances

    total = 11  # odd on purpose: the ranks own 6 and 5 instances
    offs_all = random_init_offsets(total, 4, seed=0)
    idx = shard_instances(total, rank, world)
    mine = offs_all[idx]
    # stand-in for the per-rank solve: a deterministic function of the instance data
    converged = float((np.abs(mine).sum(axis=(1, 2)) > 0).sum())
    t = torch.tensor([0.5 + rank], dtype=torch.float64)
    c = torch.tensor([converged, float(mine.sum())], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    # results travel back with their global instance index (padded to the largest shard: all_gather needs equal shapes)
    per = (total + world - 1) // world
    pad = np.full((per, 1 + 12), np.nan)
    pad[: len(idx), 0], pad[: len(idx), 1:] = idx, mine.reshape(len(idx), -1)
    gathered = [torch.zeros(per, 13, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.as_tensor(pad))
    if rank == 0:
        out.put((t.item(), c.tolist(), torch.cat(gathered).numpy(), offs_all))
    dist.destroy_process_group()


def test_instances_shard_without_overlap_and_reduce_like_bench():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    t, c, gathered, offs_all = q.get()
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert t == 1.5  # max over ranks
    assert c[0] == 11 and np.isclose(c[1], offs_all.sum())
    rows = gathered[~np.isnan(gathered[:, 0])]
    assert sorted(rows[:, 0].astype(int).tolist()) == list(range(11))  # every instance owned by exactly one rank
    back = np.zeros_like(offs_all)
    back[rows[:, 0].astype(int)] = rows[:, 1:].reshape(-1, 4, 3)
    assert np.array_equal(back, offs_all)


def test_longest_expected_first_order():
    """obca_set_order: the queue hands out the instances by decreasing predicted cost, ties in index order (deterministic)."""
    import torch

    cost = torch.tensor([5, 9, 9, 1, 7])
    order = torch.argsort(cost, descending=True, stable=True).tolist()
    assert order == [1, 2, 4, 0, 3]
