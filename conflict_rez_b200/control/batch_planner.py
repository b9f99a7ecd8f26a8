"""Batched centralised planning pipeline (the data flow of ``MultiVehiclePlanner``).

Reference flow (confrez/control/multi_vehicle_planner.py:659-667): ``solve_single_problems`` (one collocation OBCA
solve per agent, :68-109) -> ``joint_dual_ws`` (:208-341) -> ``solve_final_problem_obca`` (:343-480).  Here every stage
runs over a batch of B independent instances (different ``init_offsets``): the per-agent solves and the joint solve go
through :class:`conflict_rez_b200.solver.ObcaSolver` (CUDA), the dual warm starts are the closed forms of
``control.warmstart`` evaluated on the device (``obca_dual_ws`` / ``obca_joint_dual_ws``).
"""
import dataclasses
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from conflict_rez_b200.control import warmstart
from conflict_rez_b200.control.scenario import kinematic_paths, build_guess, build_problem, pose_guess
from conflict_rez_b200.problem import CollocationGuess, CollocationProblem
from conflict_rez_b200.solver import BatchResult, ObcaSolver, SolveOptions


def random_init_offsets(batch: int, n_vehicles: int, seed: int = 0) -> np.ndarray:
    """(B, V, 3) offsets: dx, dy ~ U(-0.15, 0.15) m, dpsi ~ U(-pi/20, pi/20)."""
    rng = np.random.default_rng(seed)
    off = rng.uniform(-1.0, 1.0, size=(batch, n_vehicles, 3))
    return off * np.array([0.15, 0.15, np.pi / 20])


def shard_instances(global_batch: int, rank: int, world: int) -> np.ndarray:
    """Indices of the instances rank ``rank`` owns: a fixed GLOBAL batch dealt out round robin (instance b -> rank b % world).
    The instances are i.i.d. draws, so every rank gets the same expected load and no data-path collective is needed
    (SURVEY.md 8e); with the contiguous block partition of round 1 one unlucky rank set the step time."""
    return np.arange(rank, global_batch, world)


def joint_guess_from_singles(prob: CollocationProblem, singles: Sequence[BatchResult]) -> CollocationGuess:
    """Joint warm start: per-agent single solutions, dt0 = mean of the agents' dt (multi_vehicle_planner.py:360),
    pair duals from the closed-form ``joint_dual_ws``."""
    B = singles[0].z.shape[0]
    V, O = prob.V, prob.O
    Mmax = int(prob.nodes.max())
    z = np.zeros((B, V, Mmax, 7))
    lam = np.zeros((B, V, Mmax, O, 4))
    mu = np.zeros((B, V, Mmax, O, 4))
    for a, r in enumerate(singles):
        M = int(prob.nodes[a])
        z[:, a, :M] = r.z[:, 0, :M]
        lam[:, a, :M] = r.lam[:, 0, :M]
        mu[:, a, :M] = r.mu[:, 0, :M]
    dt0 = np.mean([r.dt for r in singles], axis=0)
    P = len(prob.pairs)
    pl = np.zeros((B, P, Mmax, 4))
    pm = np.zeros((B, P, Mmax, 4))
    ps = np.zeros((B, P, Mmax, 2))
    for q, (a, b) in enumerate(prob.pairs):
        m = int(min(prob.nodes[a], prob.nodes[b]))
        za, zb = z[:, a, :m], z[:, b, :m]
        pl[:, q, :m], pm[:, q, :m], ps[:, q, :m] = warmstart.joint_dual_ws_rect(
            za[..., 0], za[..., 1], za[..., 2], zb[..., 0], zb[..., 1], zb[..., 2], prob.body_G, prob.body_g
        )
    return CollocationGuess(z, lam, mu, dt0, pl, pm, ps)


@dataclass
class JointPlan:
    problem: CollocationProblem
    guess: CollocationGuess
    singles: List[BatchResult]
    result: Optional[BatchResult] = None
    dev_guess: Optional[dict] = None  # the joint warm start as device tensors (ObcaSolver.set_inputs format)
    timing: Optional[dict] = None
    solver: Optional[ObcaSolver] = None  # the joint handle (kept open for the caller; close() when done)


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def tube_following_ws(p1: CollocationProblem, d: dict, options, device, lib=None):
    """State warm start on the device: the collocation problem of ``p1`` with the obstacles removed (dynamics, tube sets,
    terminal conditions, the same cost), started from the kinematic guess in ``d`` (device tensors pose, z, dt).
    Returns (z, dt, status); instances whose solve broke down keep their guess."""
    p0 = dataclasses.replace(p1, obs_A=np.zeros((0, 4, 2)), obs_b=np.zeros((0, 4)))
    s0 = ObcaSolver(p0, options, device=device, lib=lib)
    B, M = s0.B, s0.Mmax
    empty = torch.zeros((B, s0.V, M, 0, 4), dtype=torch.float64, device=d["z"].device)
    s0.set_inputs({"pose": d["pose"], "z": d["z"], "dt": d["dt"], "lam": empty, "mu": empty})
    s0.run()
    st, _, _ = s0.fetch_stats()
    sol = s0.fetch_solution()
    ok = st >= -2
    z = torch.where(ok.view(B, 1, 1, 1), sol["z"], d["z"])
    dt = torch.where(ok, sol["dt"], d["dt"])
    s0.close()
    return z, dt, st.cpu().numpy()


def euler_state_ws(p1: CollocationProblem, kin: np.ndarray, N_ws: int, dt_ws: float, device, lib=None):
    """The reference's own state warm start for a batch (``Vehicle.state_ws`` -> ``interp_ws_for_collocation``, vehicle.py:99-231 and
    :298-358): the Euler-discretised tube-following NLP solved by ``ObcaStateWsSolver`` from the kinematic guess ``kin`` (B,T,7),
    then interpolated linearly onto the collocation nodes on the device.  Returns (z (B,1,M,7), dt (B), status); instances whose
    solve failed keep their guess."""
    from conflict_rez_b200.solver import ObcaStateWsSolver, StateWsProblem, TrajectoryOps

    B, T = kin.shape[:2]
    hd = p1.final_heading[0]
    sp = StateWsProblem(tube_A=p1.tube_A[0], tube_b=p1.tube_b[0], N=N_ws, dt=dt_ws, final_heading=None if np.isnan(hd) else float(hd), shrink_tube=p1.shrink_tube,
                        wb=p1.wb, region=p1.region, limits=p1.limits, batch=B)
    assert sp.nodes == T, (sp.nodes, T)
    sv = ObcaStateWsSolver(sp, SolveOptions(tol=1e-2, constr_viol_tol=1e-2, max_iter=500), device=device, lib=lib)
    cur = np.concatenate([p1.init_pose[:, 0], np.zeros((B, 2))], axis=1)
    g = sv.upload(CollocationGuess(kin[:, None], np.zeros((B, 1, T, 0, 4)), np.zeros((B, 1, T, 0, 4)), np.zeros(B)))
    sv.set_params(sv.upload_params(cur, np.zeros((B, T, 3)), None))
    sv.set_inputs(g)
    sv.run()
    st, _, _ = sv.fetch_stats()
    sol = sv.fetch_solution()["z"]
    zt = torch.where((st >= 0).view(B, 1, 1, 1), sol, g["z"])[:, 0]  # (B,T,7)
    zt[:, -1, 5:] = zt[:, -2, 5:]  # inputs padded by their last value (vehicle.py:227-229)
    ops = TrajectoryOps(device, lib=sv.lib)
    N = int(p1.N[0])
    t = torch.arange(T, dtype=torch.float64, device=zt.device) * dt_ws
    z = ops.interp_ws(zt, t, N)[:, None]
    dt = torch.full((B,), (T - 1) * dt_ws / N, dtype=torch.float64, device=zt.device)
    sv.close()
    return z.contiguous(), dt, st.cpu().numpy()


def prepare_joint_batch(
    rl_file_name: str,
    agents: Sequence[str],
    init_offsets: np.ndarray,
    options: Optional[SolveOptions] = None,
    device="cuda:0",
    lib=None,
    final_headings: Optional[Dict[str, float]] = None,
    state_ws: str = "euler",
    **problem_kwargs,
) -> JointPlan:
    """Everything up to (not including) the joint solve, device resident, per agent the reference's
    chain state_ws -> dual_ws -> single OBCA solve (multi_vehicle_planner.py:68-109): vectorised pose guess on the host,
    obstacle-free tube-following solve (``tube_following_ws``), obstacle duals (``obca_dual_ws``), batched single-vehicle
    solve; then the joint warm start is assembled in HBM with the pair duals from ``obca_joint_dual_ws``.
    ``state_ws``: "euler" (default) = the reference's own Euler-discretised state warm start followed by ``interp_ws_for_collocation``
    (``euler_state_ws``); "collocation" = the tube-following problem in collocation form (``tube_following_ws``, the round-1 pipeline:
    same joint plans, 1.5 s instead of 1.1 s per 512 instances, profiles/r02k_ws_pipeline_euler_vs_collocation.txt)."""
    init_offsets = np.asarray(init_offsets, dtype=float)
    tm = {}
    t0 = time.perf_counter()
    prob = build_problem(rl_file_name, list(agents), init_offsets=init_offsets, final_headings=final_headings, **problem_kwargs)
    kin_all = None
    if state_ws == "euler":  # the Euler warm start takes the kinematic guess on its own uniform grid: no collocation resampling needed on the host
        kin_all = kinematic_paths(prob, rl_file_name, list(agents))
        z0 = dts = None
    else:
        z0, dts = pose_guess(prob, rl_file_name, list(agents))
    tm["host_pose_guess_s"] = time.perf_counter() - t0
    B, V, O = init_offsets.shape[0], prob.V, prob.O
    Mmax = int(prob.nodes.max())
    joint = ObcaSolver(prob, options, device=device, lib=lib)
    dev = joint.device
    t0 = time.perf_counter()
    zj = torch.zeros((B, V, Mmax, 7), dtype=torch.float64, device=dev)
    lamj = torch.zeros((B, V, Mmax, O, 4), dtype=torch.float64, device=dev)
    muj = torch.zeros_like(lamj)
    dt_sum = torch.zeros(B, dtype=torch.float64, device=dev)
    singles = []
    ws_fail = 0
    for ia, agent in enumerate(agents):
        p1 = build_problem(rl_file_name, [agent], init_offsets=init_offsets[:, ia : ia + 1], final_headings=final_headings, **problem_kwargs)
        M = int(p1.nodes[0])
        sv = ObcaSolver(p1, options, device=device, lib=lib)
        d = {"pose": sv._to_dev(p1.init_pose, (B, 1, 3))}
        if z0 is not None:
            d["z"], d["dt"] = sv._to_dev(z0[:, ia : ia + 1, :M], (B, 1, M, 7)), sv._to_dev(dts[:, ia], (B,))
        # stage 1 (the role of Vehicle.state_ws, vehicle.py:99-231): the same tube-following problem without obstacles gives a
        # dynamically feasible trajectory inside the tube sets; the OBCA solve then starts from it
        if state_ws == "euler":
            z1, dt1, st1 = euler_state_ws(p1, kin_all[ia], 30, 0.1, device, lib)
        else:
            z1, dt1, st1 = tube_following_ws(p1, d, options, device, lib)
        ws_fail += int((st1 < 0).sum())
        d["z"], d["dt"] = z1, dt1
        d["lam"], d["mu"] = sv.dual_ws(d["z"])
        sv.set_inputs(d)
        sv.run()
        st, it, dbl = sv.fetch_stats()
        sol = sv.fetch_solution()
        # a solve that stopped on the iteration limit or in the line search still is the better warm start (measured: the
        # joint solve needs <= 123 iterations from it, up to 259 from the raw guess); only unusable iterates are replaced
        ok = st >= -2
        pick = lambda a, b: torch.where(ok.view((B,) + (1,) * (a.dim() - 1)), a, b)
        zj[:, ia, :M] = pick(sol["z"], d["z"])[:, 0]
        lamj[:, ia, :M] = pick(sol["lam"], d["lam"])[:, 0]
        muj[:, ia, :M] = pick(sol["mu"], d["mu"])[:, 0]
        dt_sum += pick(sol["dt"], d["dt"])
        cpu = lambda t: t.cpu().numpy()
        singles.append(BatchResult(cpu(st), cpu(it), cpu(dbl[0]), cpu(dbl[1]), cpu(dbl[2]), cpu(dbl[3]), cpu(dbl[4]), cpu(sol["z"]), None, None, cpu(sol["dt"]), None, None, None))
        sv.close()
    dg = {"pose": joint._to_dev(prob.init_pose, (B, V, 3)), "z": zj, "lam": lamj, "mu": muj, "dt": dt_sum / V}  # dt0: multi_vehicle_planner.py:360
    if joint.P:
        dg["pl"], dg["pm"], dg["ps"] = joint.joint_dual_ws(zj)
    _sync(dev)
    tm["device_warm_start_s"] = time.perf_counter() - t0
    host = lambda k: dg[k].cpu().numpy() if k in dg else None
    guess = CollocationGuess(host("z"), host("lam"), host("mu"), host("dt"), host("pl"), host("pm"), host("ps"))
    tm["state_ws_failures"] = ws_fail
    return JointPlan(prob, guess, singles, dev_guess=dg, timing=tm, solver=joint)


def solve_joint_batch(rl_file_name, agents, init_offsets, options=None, device="cuda:0", lib=None, **kw) -> JointPlan:
    plan = prepare_joint_batch(rl_file_name, agents, init_offsets, options, device, lib, **kw)
    t0 = time.perf_counter()
    plan.result = plan.solver.solve(plan.guess)
    plan.timing["joint_solve_s"] = time.perf_counter() - t0
    plan.solver.close()
    plan.solver = None
    return plan


def replicate_problem(prob: CollocationProblem, copies: int, shift=(35.0, 0.0)) -> CollocationProblem:
    """``copies`` parking lots side by side (copy c translated by c * shift) in ONE joint problem of ``copies * V`` vehicles and
    ``copies * O`` obstacles: the synthetic scenario of the scaling sweep beyond the four agents of the RL environment (SURVEY.md 8d,
    config 5: 2-8 vehicles).  Every pair of vehicles keeps its collision-avoidance block, also across lots (far apart: inactive)."""
    s = np.asarray(shift, dtype=float)
    tile = lambda a: np.concatenate([a] * copies, axis=0)
    obs_b = np.concatenate([prob.obs_b + c * (prob.obs_A @ s) for c in range(copies)], axis=0)
    tube_b = np.concatenate([prob.tube_b + c * (prob.tube_A @ s) for c in range(copies)], axis=0)
    off = np.concatenate([np.array([c * s[0], c * s[1], 0.0])[None].repeat(prob.V, 0) for c in range(copies)], axis=0)
    init = np.concatenate([prob.init_pose] * copies, axis=-2) + off
    region = prob.region.copy()
    region[1] += (copies - 1) * max(s[0], 0.0)
    region[3] += (copies - 1) * max(s[1], 0.0)
    return dataclasses.replace(prob, n_sets=tile(prob.n_sets), obs_A=tile(prob.obs_A), obs_b=obs_b, tube_A=tile(prob.tube_A), tube_b=tube_b,
                               init_pose=init, final_heading=tile(prob.final_heading), region=region)


def prepare_replicated_batch(rl_file_name: str, agents: Sequence[str], copies: int, init_offsets: np.ndarray, options: Optional[SolveOptions] = None,
                             device="cuda:0", lib=None, shift=(35.0, 0.0), **kw) -> JointPlan:
    """Warm start + joint handle for ``copies`` replicated lots (``copies * len(agents)`` vehicles per instance):
    ``init_offsets`` (B, copies * V, 3).  Each lot runs the reference chain (``prepare_joint_batch``) in its own frame; the joint
    guess concatenates the single-vehicle solutions (translated), the duals of all ``copies * O`` obstacles and of all pairs come
    from the closed-form device warm starts."""
    V = len(agents)
    init_offsets = np.asarray(init_offsets, dtype=float)
    assert init_offsets.shape[1] == copies * V
    t0 = time.perf_counter()
    lots = [prepare_joint_batch(rl_file_name, agents, init_offsets[:, c * V:(c + 1) * V], options, device, lib, **kw) for c in range(copies)]
    for p in lots:
        p.solver.close()
    base = lots[0].problem
    prob = replicate_problem(base, copies, shift)
    prob.init_pose = np.concatenate([lots[c].problem.init_pose + np.array([c * shift[0], c * shift[1], 0.0]) for c in range(copies)], axis=1)
    joint = ObcaSolver(prob, options, device=device, lib=lib)
    zs = []
    for c, p in enumerate(lots):
        z = p.dev_guess["z"].clone()
        z[..., 0] += c * shift[0]
        z[..., 1] += c * shift[1]
        zs.append(z)
    zj = torch.cat(zs, dim=1).contiguous()
    dg = {"pose": joint._to_dev(prob.init_pose, (joint.B, joint.V, 3)), "z": zj, "dt": torch.stack([p.dev_guess["dt"] for p in lots]).mean(0)}
    dg["lam"], dg["mu"] = joint.dual_ws(zj)
    dg["pl"], dg["pm"], dg["ps"] = joint.joint_dual_ws(zj)
    _sync(joint.device)
    host = lambda k: dg[k].cpu().numpy()
    guess = CollocationGuess(host("z"), host("lam"), host("mu"), host("dt"), host("pl"), host("pm"), host("ps"))
    singles = [r for p in lots for r in p.singles]
    return JointPlan(prob, guess, singles, dev_guess=dg, timing={"warm_start_s": time.perf_counter() - t0}, solver=joint)
