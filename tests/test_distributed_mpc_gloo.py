"""Distributed MPC over ``torch.distributed`` (world_size 2, gloo, host emulation of the kernels): two vehicles per rank,
one all-gather of the predicted poses per control step.  The closed-loop trajectories must equal the single-process
``MultiDistributedFollower`` (Jacobi snapshot, vehicle_follower.py:636-637) exactly."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

AGENTS = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
HEADS = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
STEPS = 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(cls, strategy_file, lib, **kw):
    from conflict_rez_b200.pytypes import VehicleState

    np.random.seed(0)  # the first dual warm start is 0.1 * rand (vehicle_follower.py:401-402)
    return cls(strategy_file, {a: True for a in AGENTS}, {a: {} for a in AGENTS}, {a: VehicleState() for a in AGENTS}, HEADS, device="cpu", lib=lib, **kw)


def _worker(rank, world, port, strategy_file, lib_path, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conflict_rez_b200 import solver
    from conflict_rez_b200.control.vehicle_follower import DistributedFollowerNode

    node = _make(DistributedFollowerNode, strategy_file, solver.load_library(lib_path))
    node.setup_multi_vehicles()
    # the dual warm start is random per vehicle: use the same draws as the single-process run (agent order)
    rng = np.random.RandomState(1)
    draws = {a: (0.1 * rng.rand(node.N, 24), 0.1 * rng.rand(node.N, 24)) for a in AGENTS}
    for v in node.vehicles:
        v.pred.l, v.pred.m = draws[v.agent][0].copy(), draws[v.agent][1].copy()
    node.solve(num_iter=STEPS)
    out.put((rank, {a: np.stack([t.x, t.y, t.psi], axis=1) for a, t in node.final_results.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_closed_loop_equals_single_process(strategy_file, emu_lib):
    from conflict_rez_b200.control.vehicle_follower import MultiDistributedFollower

    lib_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "host_emu", "libobca_hostemu.so")
    ref = _make(MultiDistributedFollower, strategy_file, emu_lib)
    ref.setup_multi_vehicles()
    rng = np.random.RandomState(1)
    for a in AGENTS:
        v = ref.vehicles[ref.agents.index(a)]
        v.pred.l, v.pred.m = 0.1 * rng.rand(v.N, 24), 0.1 * rng.rand(v.N, 24)
    ref.solve(num_iter=STEPS)
    want = {a: np.stack([t.x, t.y, t.psi], axis=1) for a, t in ref.final_results.items()}

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, strategy_file, lib_path, q)) for r in range(world)]
    [p.start() for p in procs]
    got = {}
    import time

    t_end = time.time() + 300
    while len(got) < len(AGENTS) and time.time() < t_end and (not q.empty() or all(p.exitcode in (None, 0) for p in procs)):
        if q.empty():
            time.sleep(0.2)
            continue
        _, part = q.get()
        got.update(part)
    [p.join(60) for p in procs]
    [p.kill() for p in procs if p.is_alive()]
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(got) == AGENTS
    for a in AGENTS:
        assert got[a].shape == (STEPS + 1, 3)
        np.testing.assert_allclose(got[a], want[a], rtol=0, atol=1e-12)
