"""Strategy file -> problem descriptor + warm start (the data ``setup_single_final_problem`` consumes).

Reference data flow: ``Vehicle.__init__`` (confrez/control/vehicle.py:46-52) reads initial pose, obstacles
and tube sets; ``main`` of each planner supplies ``init_offsets`` / ``final_headings``
(multi_vehicle_planner.py:636-649).
"""
from typing import Dict, List, Optional, Sequence

import numpy as np

from conflict_rez_b200.control.compute_sets import (
    compute_initial_states,
    compute_obstacles,
    compute_sets,
    interp_along_sets,
)
from conflict_rez_b200.control import warmstart
from conflict_rez_b200.problem import CollocationGuess, CollocationProblem
from conflict_rez_b200.vehicle_types import VehicleBody, VehicleConfig
from conflict_rez_b200.obstacle_types import GeofenceRegion

DEFAULT_FINAL_HEADINGS = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}


def build_problem(
    rl_file_name: str,
    agents: Sequence[str],
    init_offsets: Optional[np.ndarray] = None,
    final_headings: Optional[Dict[str, float]] = None,
    K: int = 5,
    n_per_set: int = 5,
    shrink_tube: float = 0.5,
    dmin: float = 0.05,
    vehicle_body: Optional[VehicleBody] = None,
    vehicle_config: Optional[VehicleConfig] = None,
    region: Optional[GeofenceRegion] = None,
    obstacles: Optional[List] = None,
    rl_tubes: Optional[Dict] = None,
) -> CollocationProblem:
    """``init_offsets``: ([B,] V, 3) offsets on (x, y, psi) added to the strategy's initial poses."""
    vb = vehicle_body or VehicleBody()
    vc = vehicle_config or VehicleConfig()
    rg = region or GeofenceRegion()
    obstacles = compute_obstacles(vb=vb) if obstacles is None else obstacles
    tubes = compute_sets(rl_file_name) if rl_tubes is None else rl_tubes
    init = compute_initial_states(rl_file_name, vb)
    V = len(agents)
    n_sets = np.array([len(tubes[a]) for a in agents])
    Smax = int(n_sets.max())
    tube_A = np.zeros((V, Smax, 2, 4, 2))
    tube_b = np.zeros((V, Smax, 2, 4))
    for ia, a in enumerate(agents):
        for q, sets in enumerate(tubes[a]):
            for ib, body in enumerate(("back", "front")):
                tube_A[ia, q, ib] = sets[body].A
                tube_b[ia, q, ib] = np.ravel(sets[body].b)
    base = np.array([[init[a].x.x, init[a].x.y, init[a].e.psi] for a in agents])
    init_pose = base if init_offsets is None else base + np.asarray(init_offsets, dtype=float)
    fh = DEFAULT_FINAL_HEADINGS if final_headings is None else final_headings
    heading = np.array([np.nan if fh.get(a) is None else float(fh[a]) for a in agents])
    return CollocationProblem(
        n_sets=n_sets,
        obs_A=np.stack([o.A for o in obstacles]),
        obs_b=np.stack([np.ravel(o.b) for o in obstacles]),
        tube_A=tube_A,
        tube_b=tube_b,
        init_pose=init_pose,
        final_heading=heading,
        body_G=np.asarray(vb.A, dtype=float),
        body_g=np.asarray(vb.b, dtype=float),
        wb=vb.wb,
        region=np.array([rg.x_min, rg.x_max, rg.y_min, rg.y_max]),
        limits=np.array([vc.v_min, vc.v_max, vc.delta_min, vc.delta_max, vc.a_min, vc.a_max, vc.w_delta_min, vc.w_delta_max], dtype=float),
        K=K,
        n_per_set=n_per_set,
        dmin=dmin,
        shrink_tube=shrink_tube,
    )


def kinematic_paths(prob: CollocationProblem, rl_file_name: str, agents: Sequence[str], N_ws: int = 30, dt_ws: float = 0.1):
    """Per agent the kinematic guess on the uniform grid of ``Vehicle.state_ws`` (N_ws samples per move): list of (B,T_a,7) arrays
    (x, y, psi, v, delta, a, w), blended towards each instance's perturbed initial pose like ``pose_guess``."""
    vb = VehicleBody()
    paths = interp_along_sets(rl_file_name, vb, N_ws)
    init = prob.init_pose if prob.batch is not None else prob.init_pose[None]
    out = []
    for ia, a in enumerate(agents):
        path0 = paths[a]
        blend = np.clip(1.0 - np.arange(len(path0)) / float(N_ws), 0.0, 1.0)[:, None]
        path = path0[None] + blend[None] * (init[:, ia] - path0[0])[:, None, :]
        kin = warmstart.kinematic_guess(path, dt_ws, prob.wb, prob.limits)
        out.append(np.stack([kin[k] for k in ("x", "y", "psi", "v", "delta", "a", "w")], axis=-1))
    return out


def pose_guess(prob: CollocationProblem, rl_file_name: str, agents: Sequence[str], N_ws: int = 30, dt_ws: float = 0.1):
    """Spline pose guess -> kinematic state guess -> Radau resampling, vectorised over the batch: z (B,V,Mmax,7), dts (B,V).

    The pose guess is shifted rigidly at t=0 towards each instance's perturbed initial pose and blended out over the
    first set move, so that every instance starts from a guess consistent with its own initial condition.
    """
    vb = VehicleBody()
    paths = interp_along_sets(rl_file_name, vb, N_ws)
    init = prob.init_pose if prob.batch is not None else prob.init_pose[None]
    B, V = init.shape[0], prob.V
    Mmax = int(prob.nodes.max())
    z = np.zeros((B, V, Mmax, 7))
    dts = np.zeros((B, V))
    for ia, a in enumerate(agents):
        path0 = paths[a]
        N, M = int(prob.N[ia]), int(prob.nodes[ia])
        blend = np.clip(1.0 - np.arange(len(path0)) / float(N_ws), 0.0, 1.0)[:, None]
        path = path0[None] + blend[None] * (init[:, ia] - path0[0])[:, None, :]  # (B,T,3)
        kin = warmstart.kinematic_guess(path, dt_ws, prob.wb, prob.limits)
        for c, k in enumerate(("x", "y", "psi", "v", "delta", "a", "w")):
            z[:, ia, :M, c] = warmstart.resample_for_collocation(kin["t"], kin[k], N, prob.K)
        dts[:, ia] = kin["t"][-1] / N
    return z, dts


def build_guess(prob: CollocationProblem, rl_file_name: str, agents: Sequence[str], N_ws: int = 30, dt_ws: float = 0.1) -> CollocationGuess:
    """``pose_guess`` + closed-form duals on the host (the batched planners compute the duals on the device instead,
    ``ObcaSolver.dual_ws`` / ``joint_dual_ws``; both implement control/warmstart.py's formulas)."""
    batched = prob.batch is not None
    z, dts = pose_guess(prob, rl_file_name, agents, N_ws, dt_ws)
    B, V, O = z.shape[0], prob.V, prob.O
    Mmax = int(prob.nodes.max())
    lam = np.zeros((B, V, Mmax, O, 4))
    mu = np.zeros((B, V, Mmax, O, 4))
    for ia in range(V):
        M = int(prob.nodes[ia])
        zz = z[:, ia, :M]
        lam[:, ia, :M], mu[:, ia, :M] = warmstart.dual_ws_rect(zz[..., 0], zz[..., 1], zz[..., 2], prob.obs_A, prob.obs_b, prob.body_G, prob.body_g)
    P = len(prob.pairs)
    pl = np.zeros((B, P, Mmax, 4))
    pm = np.zeros((B, P, Mmax, 4))
    ps = np.zeros((B, P, Mmax, 2))
    for q, (a, b_) in enumerate(prob.pairs):
        m = int(min(prob.nodes[a], prob.nodes[b_]))
        za, zb = z[:, a, :m], z[:, b_, :m]
        l_, m_, s_ = warmstart.joint_dual_ws_rect(za[..., 0], za[..., 1], za[..., 2], zb[..., 0], zb[..., 1], zb[..., 2], prob.body_G, prob.body_g)
        pl[:, q, :m], pm[:, q, :m], ps[:, q, :m] = l_, m_, s_
    dt0 = dts.mean(axis=1)  # multi_vehicle_planner.py:360
    g = CollocationGuess(z, lam, mu, dt0, pl, pm, ps)
    return g if batched else g.instance(0)


def random_obstacles(rl_file_name: str, n_extra: int, seed: int = 0, clearance: float = 0.6):
    """The six parking-row rectangles plus ``n_extra`` seeded axis-aligned rectangles (obstacle
    count sweep).  A candidate is rejected when it comes closer than ``clearance`` to any tube set of any agent, so the
    strategy stays feasible."""
    from conflict_rez_b200.control.compute_sets import _rect

    rng = np.random.default_rng(seed)
    base = compute_obstacles()
    tubes = compute_sets(rl_file_name)
    boxes = []
    for sets in tubes.values():
        for s in sets:
            for body in ("back", "front"):
                V = np.asarray(s[body].V)
                boxes.append((V[:, 0].min(), V[:, 0].max(), V[:, 1].min(), V[:, 1].max()))
    rg = GeofenceRegion()
    out, tries = [], 0
    while len(out) < n_extra and tries < 10000:
        tries += 1
        w, h = rng.uniform(0.5, 2.0, size=2)
        x0, y0 = rng.uniform(rg.x_min, rg.x_max - w), rng.uniform(rg.y_min, rg.y_max - h)
        if any(x0 - clearance < b[1] and x0 + w + clearance > b[0] and y0 - clearance < b[3] and y0 + h + clearance > b[2] for b in boxes):
            continue
        out.append(_rect(x0, x0 + w, y0, y0 + h))
    if len(out) < n_extra:
        raise RuntimeError("could not place %d obstacles" % n_extra)
    return base + out
