"""Batched centralised planning pipeline (the data flow of ``MultiVehiclePlanner``).

Reference flow (confrez/control/multi_vehicle_planner.py:659-667): ``solve_single_problems`` (one collocation OBCA
solve per agent, :68-109) -> ``joint_dual_ws`` (:208-341) -> ``solve_final_problem_obca`` (:343-480).  Here every stage
runs over a batch of B independent instances (different ``init_offsets``): the per-agent solves and the joint solve go
through :class:`conflict_rez_b200.solver.ObcaSolver` (CUDA), the pair-dual warm start is the closed form of
``control.warmstart``.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from conflict_rez_b200.control import warmstart
from conflict_rez_b200.control.scenario import build_guess, build_problem
from conflict_rez_b200.problem import CollocationGuess, CollocationProblem
from conflict_rez_b200.solver import BatchResult, ObcaSolver, SolveOptions


def random_init_offsets(batch: int, n_vehicles: int, seed: int = 0) -> np.ndarray:
    """(B, V, 3) offsets: dx, dy ~ U(-0.15, 0.15) m, dpsi ~ U(-pi/20, pi/20)."""
    rng = np.random.default_rng(seed)
    off = rng.uniform(-1.0, 1.0, size=(batch, n_vehicles, 3))
    return off * np.array([0.15, 0.15, np.pi / 20])


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


def prepare_joint_batch(
    rl_file_name: str,
    agents: Sequence[str],
    init_offsets: np.ndarray,
    options: Optional[SolveOptions] = None,
    device="cuda:0",
    lib=None,
    final_headings: Optional[Dict[str, float]] = None,
    **problem_kwargs,
) -> JointPlan:
    """Everything up to (not including) the joint solve: batched single-vehicle solves + joint warm start."""
    init_offsets = np.asarray(init_offsets, dtype=float)
    singles = []
    for ia, agent in enumerate(agents):
        p1 = build_problem(rl_file_name, [agent], init_offsets=init_offsets[:, ia : ia + 1], final_headings=final_headings, **problem_kwargs)
        g1 = build_guess(p1, rl_file_name, [agent])
        sv = ObcaSolver(p1, options, device=device, lib=lib)
        singles.append(sv.solve(g1))
        sv.close()
    prob = build_problem(rl_file_name, list(agents), init_offsets=init_offsets, final_headings=final_headings, **problem_kwargs)
    return JointPlan(prob, joint_guess_from_singles(prob, singles), singles)


def solve_joint_batch(rl_file_name, agents, init_offsets, options=None, device="cuda:0", lib=None, **kw) -> JointPlan:
    plan = prepare_joint_batch(rl_file_name, agents, init_offsets, options, device, lib, **kw)
    sv = ObcaSolver(plan.problem, options, device=device, lib=lib)
    plan.result = sv.solve(plan.guess)
    sv.close()
    return plan
