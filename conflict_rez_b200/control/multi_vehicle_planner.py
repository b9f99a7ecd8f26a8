"""``MultiVehiclePlanner`` -- centralised conflict resolution with the reference's method surface
(confrez/control/multi_vehicle_planner.py:25-480).

``solve_single_problems`` runs the single-vehicle pipeline per agent, ``joint_dual_ws`` warm-starts the pair duals
(closed form of multi_vehicle_planner.py:208-341), ``solve_final_problem_obca`` assembles the joint NLP (all vehicles, one
shared ``dt``, pair blocks for ``i < N_min``) and solves it on the GPU.  ``solve_final_problem_circles`` is not provided:
the reference's version raises ``TypeError`` before solving.
"""
from itertools import combinations
from typing import Dict, Tuple

import numpy as np

from conflict_rez_b200.control import warmstart
from conflict_rez_b200.control.compute_sets import compute_obstacles, compute_sets
from conflict_rez_b200.control.vehicle import JointProblem, Vehicle
from conflict_rez_b200.obstacle_types import GeofenceRegion
from conflict_rez_b200.pytypes import VehiclePrediction, VehicleState
from conflict_rez_b200.vehicle_types import VehicleBody, VehicleConfig


class MultiVehiclePlanner(object):
    def __init__(
        self,
        rl_file_name: str,
        ws_config: Dict[str, bool],
        colors: Dict[str, Tuple[float, float, float]],
        init_offsets: Dict[str, VehicleState],
        final_headings: Dict[str, float],
        vehicle_body: VehicleBody = VehicleBody(),
        vehicle_config: VehicleConfig = VehicleConfig(),
        region: GeofenceRegion = GeofenceRegion(),
        device="cuda:0",
    ) -> None:
        self.rl_file_name, self.ws_config, self.colors = rl_file_name, ws_config, colors
        self.init_offsets, self.final_headings = init_offsets, final_headings
        self.vehicle_body, self.vehicle_config, self.region, self.device = vehicle_body, vehicle_config, region, device
        self.agents = sorted(self.ws_config.keys())
        self.agent_pairs = list(combinations(self.agents, 2))
        self.rl_tubes = compute_sets(self.rl_file_name)
        self.obstacles = compute_obstacles()
        self.vehicles = {
            agent: Vehicle(rl_file_name=self.rl_file_name, agent=agent, color=self.colors[agent], vehicle_config=self.vehicle_config,
                           vehicle_body=self.vehicle_body, region=self.region, device=device)
            for agent in self.agents
        }

    def solve_single_problems(self, N: int = 30, K: int = 5, N_per_set: int = 5, dt: float = 0.1, shrink_tube: float = 0.5, dmin: float = 0.05):
        self.single_results = {agent: VehiclePrediction() for agent in self.agents}
        for agent in self.agents:
            vehicle = self.vehicles[agent]
            zu0 = vehicle.state_ws(N=N, dt=dt, init_offset=self.init_offsets[agent], final_heading=self.final_headings[agent],
                                   shrink_tube=shrink_tube, spline_ws=self.ws_config[agent])
            zu0 = vehicle.dual_ws(zu0=zu0)
            zu0 = vehicle.interp_ws_for_collocation(zu0=zu0, K=K, N_per_set=N_per_set)
            vehicle.setup_single_final_problem(zu0=zu0, init_offset=self.init_offsets[agent], final_heading=self.final_headings[agent],
                                               K=K, N_per_set=N_per_set, shrink_tube=shrink_tube, dmin=dmin)
            sol = vehicle.solve_single_final_problem()
            self.single_results[agent] = vehicle.get_solution(sol=sol)

    def joint_dual_ws(self, K=5, verbose=0):
        """Pair duals at the single-vehicle solutions: lambda_ij, lambda_ji in R^4_+, s in R^2 per (pair, node)."""
        G, g = np.asarray(self.vehicle_body.A, float), np.asarray(self.vehicle_body.b, float)
        self.joint_l0, self.joint_s0 = {a: {} for a in self.agents}, {}
        for agent, other in self.agent_pairs:
            ra, rb = self.single_results[agent], self.single_results[other]
            m = min(len(ra.x), len(rb.x))
            lam, mu, s = warmstart.joint_dual_ws_rect(ra.x[:m], ra.y[:m], ra.psi[:m], rb.x[:m], rb.y[:m], rb.psi[:m], G, g)
            N_min = m // (K + 1)
            shape = lambda arr: [[arr[i * (K + 1) + k] for k in range(K + 1)] for i in range(N_min)]
            self.joint_l0[agent][other], self.joint_l0[other][agent], self.joint_s0[(agent, other)] = shape(lam), shape(mu), shape(s)

    def solve_final_problem_obca(self, K: int = 5, N_per_set: int = 5, shrink_tube: float = 0.5, dmin: float = 0.05, interp_dt: float = None):
        self.joint_dual_ws(K=K)
        dt0 = float(np.mean([self.single_results[agent].dt for agent in self.agents]))
        opti = JointProblem(self.obstacles, self.vehicle_body, self.vehicle_config, self.region)
        for agent in self.agents:
            self.vehicles[agent].setup_single_final_problem(
                zu0=self.single_results[agent], init_offset=self.init_offsets[agent], final_heading=self.final_headings[agent],
                opti=opti, dt=dt0, K=K, N_per_set=N_per_set, dmin=dmin, shrink_tube=shrink_tube)
        Mmax = max(self.vehicles[a].N for a in self.agents) * (K + 1)
        P = len(self.agent_pairs)
        pl, pm, ps = np.zeros((P, Mmax, 4)), np.zeros((P, Mmax, 4)), np.zeros((P, Mmax, 2))
        for q, (agent, other) in enumerate(self.agent_pairs):
            flat = lambda nested: np.array([v for row in nested for v in row])
            m = len(self.joint_l0[agent][other]) * (K + 1)
            pl[q, :m], pm[q, :m], ps[q, :m] = flat(self.joint_l0[agent][other]), flat(self.joint_l0[other][agent]), flat(self.joint_s0[(agent, other)])
        opti.set_pair_initial(pl, pm, ps)
        first = self.vehicles[self.agents[0]]
        sol = opti.solve(first.solve_options, self.device, getattr(first, "_lib", None))
        self.final_sol = sol
        N_max = max(self.vehicles[agent].N for agent in self.agents)
        dt = float(sol.result.dt[0])
        if interp_dt is None:
            final_t = np.linspace(0, N_max * dt, N_max * (K + 1) + 1, endpoint=True)
        else:
            final_t = np.arange(0, N_max * dt, interp_dt)
        self.final_results = {agent: VehiclePrediction() for agent in self.agents}
        for agent in self.agents:
            self.vehicles[agent].get_solution(sol=sol)
            self.final_results[agent] = self.vehicles[agent].interpolate_states(final_t)

    def solve_final_problem_circles(self, K: int = 5, N_per_set: int = 5, shrink_tube: float = 0.5, dmin: float = 0.05, d_buffer: float = 0.2):
        """The circle-approximation variant of the joint problem (multi_vehicle_planner.py:111-206, rect2circles.py:13-37) is not
        provided: the reference method cannot run as shipped (``self.agents - {agent}`` on a list raises ``TypeError`` at :150; it also
        omits ``final_heading`` and only constrains the nodes ``k < K``), nothing in the reference calls it, and the OBCA formulation
        (``solve_final_problem_obca``) is the one every experiment uses.  Stated explicitly instead of failing with ``AttributeError``.
        The variant's geometry -- the circle cover ``v2c`` / ``v2c_ca`` and the values of its pair rows -- is in ``control/rect2circles.py``."""
        raise NotImplementedError("solve_final_problem_circles: the reference variant is unusable as shipped (SURVEY.md App. B-1); use solve_final_problem_obca")
