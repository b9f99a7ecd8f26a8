"""Scenario geometry and strategy -> tube sets, mirroring ``confrez/control/compute_sets.py``.

Reference: compute_sets.py:27-139 (compute_sets: one L x L square per body half
per strategy step), :142-164 (convert_rl_states), :167-240 (interp_along_sets),
:243-256 (compute_initial_states), :259-330 (compute_obstacles: six rectangles).
"""
import pickle
from typing import Dict, List, Tuple

import numpy as np

from conflict_rez_b200.polytope import Polytope
from conflict_rez_b200.control.bezier import BezierPlanner
from conflict_rez_b200.control.utils import pi_2_pi
from conflict_rez_b200.pytypes import VehicleState
from conflict_rez_b200.vehicle_types import VehicleBody


def _load(file_name: str):
    with open(file_name + ".pkl", "rb") as f:
        return pickle.load(f)


def compute_sets(file_name: str, L=2.5) -> Dict[str, List[Dict[str, Polytope]]]:
    history = _load(file_name)
    square = [[0, 0], [0, L], [L, 0], [L, L]]
    rl_sets = {agent: [] for agent in history}
    for agent in rl_sets:
        for state in history[agent]:
            rl_sets[agent].append({body: Polytope(square) + np.array(state[body]) * L for body in ("front", "back")})
    return rl_sets


def convert_rl_states(states: Dict[str, Tuple[int, int]], vehicle_body: VehicleBody, L: float = 2.5) -> VehicleState:
    vehicle_state = VehicleState()
    front, back = states["front"], states["back"]
    direction = (front[0] - back[0], front[1] - back[1])
    psi = np.arctan2(direction[1], direction[0])
    vehicle_state.e.psi = psi
    if direction[1] == 0:
        center = np.array([max(front[0], back[0]) * L, (front[1] + 0.5) * L])
    elif direction[0] == 0:
        center = np.array([(front[0] + 0.5) * L, max(front[1], back[1]) * L])
    else:
        center = np.array([max(front[0], back[0]) * L, max(front[1], back[1]) * L])
    wb = vehicle_body.wb
    vehicle_state.x.x = center[0] - wb / 2 * np.cos(psi)
    vehicle_state.x.y = center[1] - wb / 2 * np.sin(psi)
    return vehicle_state


def interp_along_sets(file_name: str, vehicle_body: VehicleBody, N: int):
    """Piecewise Bezier guess of (x, y, psi) through the strategy poses: N samples per move + the final pose."""
    history = _load(file_name)
    path = {agent: [] for agent in history}
    planner = BezierPlanner(offset=2.5)
    for agent in history:
        steps = history[agent]
        for i in range(len(steps) - 1):
            s0 = convert_rl_states(steps[i], vehicle_body)
            s1 = convert_rl_states(steps[i + 1], vehicle_body)
            hold = np.tile([s0.x.x, s0.x.y, s0.e.psi], (N, 1))
            if steps[i + 1] == steps[i]:
                seg = hold
            elif s0.e.psi == s1.e.psi:
                seg = hold
                seg[:, 0] = np.linspace(s0.x.x, s1.x.x, N, endpoint=False)
                seg[:, 1] = np.linspace(s0.x.y, s1.x.y, N, endpoint=False)
            else:
                angle_offset = np.pi if steps[i + 1]["front"] == steps[i]["back"] else 0
                s0.e.psi = pi_2_pi(s0.e.psi + angle_offset)
                s1.e.psi = pi_2_pi(s1.e.psi + angle_offset)
                seg = planner.interpolate(start_state=s0, end_state=s1, N=N)
                seg[:, 2] -= angle_offset
            path[agent].append(seg)
        sf = convert_rl_states(steps[-1], vehicle_body)
        path[agent].append(np.array([[sf.x.x, sf.x.y, sf.e.psi]]))
        path[agent] = np.vstack(path[agent])
        path[agent][:, 2] = np.unwrap(path[agent][:, 2])
    return path


def compute_initial_states(file_name: str, vehicle_body: VehicleBody, L=2.5) -> Dict[str, VehicleState]:
    history = _load(file_name)
    return {agent: convert_rl_states(history[agent][0], vehicle_body) for agent in history}


def _rect(xmin, xmax, ymin, ymax) -> Polytope:
    return Polytope([[xmin, ymin], [xmin, ymax], [xmax, ymax], [xmax, ymin]])


def compute_obstacles(L: float = 2.5, vb: VehicleBody = VehicleBody()) -> List[Polytope]:
    """The six parking-row rectangles (compute_sets.py:264-328)."""
    hw = vb.w / 2
    return [
        _rect(1.5 * L - hw, 5.5 * L + hw, 3 * L, 5.5 * L),  # bottom left
        _rect(7.5 * L - hw, 7.5 * L + hw, 3 * L, 5.5 * L),  # bottom centre
        _rect(9.5 * L - hw, 12.5 * L + hw, 3 * L, 5.5 * L),  # bottom right
        _rect(1.5 * L - hw, 5.5 * L + hw, 8.5 * L, 11 * L),  # top left
        _rect(7.5 * L - hw, 8.5 * L + hw, 8.5 * L, 11 * L),  # top centre
        _rect(10.5 * L - hw, 12.5 * L + hw, 8.5 * L, 11 * L),  # top right
    ]
