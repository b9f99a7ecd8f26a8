"""Synthetic grid "strategy" generator replacing the un-shipped RL pickle.

The reference reads ``<rl_file_name>.pkl`` = ``Dict[agent, List[{"front": (gx, gy),
"back": (gx, gy)}]]`` written by confrez/rl/record_states_history.py:10-31 from
a pretrained DQN (not shipped).  This module produces files with the *same
schema* by running a seeded prioritized space-time search on exactly the
grid/move/wall rules of confrez/rl/pklot_env.py:131-139 (actions), :226-282
(walls), :300-356 (move), :369-387 (collision), from the configured starts to
goals (:141-158).  Lists are ragged: an agent stops being appended once it
reaches its goal (:617,674-679).
"""
from collections import deque
from itertools import product
import pickle
from typing import Dict, List, Tuple

import numpy as np

N_CENTER, N_EDGE = 8, 3
N_TOTAL = N_CENTER + 2 * N_EDGE

AGENT_CONFIGS = [
    {"init_state": {"front": (6, 8), "back": (6, 7)}, "goal": {"front": (12, 6), "back": (11, 6)}},
    {"init_state": {"front": (8, 7), "back": (9, 7)}, "goal": {"front": (6, 3), "back": (6, 4)}},
    {"init_state": {"front": (6, 5), "back": (6, 4)}, "goal": {"front": (1, 7), "back": (2, 7)}},
    {"init_state": {"front": (5, 6), "back": (4, 6)}, "goal": {"front": (6, 10), "back": (6, 9)}},
]

ACTIONS = {0: (0, 0.0), 1: (1, -np.pi / 4), 2: (1, 0.0), 3: (1, np.pi / 4), 4: (-1, -np.pi / 4), 5: (-1, 0.0), 6: (-1, np.pi / 4)}


def wall_cells() -> set:
    walls = set()
    for x, y in product(range(N_TOTAL), range(N_TOTAL - N_EDGE, N_TOTAL)):
        walls.add((x, y))
    for x, y in product(range(N_TOTAL), range(N_EDGE)):
        walls.add((x, y))
    for x, y in product(range(N_EDGE), range(N_EDGE, N_EDGE + N_CENTER)):
        walls.add((x, y))
    for x, y in product(range(1, N_EDGE), range(N_EDGE + 3, N_EDGE + 5)):
        walls.discard((x, y))
    for x, y in product(range(N_EDGE + N_CENTER, N_TOTAL), range(N_EDGE, N_EDGE + N_CENTER)):
        walls.add((x, y))
    for x, y in product(range(N_EDGE + N_CENTER, N_EDGE + N_CENTER + 2), range(N_EDGE + 3, N_EDGE + 5)):
        walls.discard((x, y))
    for i in [3, 4, 5, 7, 8, 10]:
        for r in (1, 2, 3):
            walls.add((i, N_EDGE + N_CENTER - r))
    for i in [3, 4, 5, 7, 9, 10]:
        for r in (0, 1, 2):
            walls.add((i, N_EDGE + r))
    return walls


def move(state, action, walls):
    """One grid move (pklot_env.py:300-356); returns the new (front, back) or None when it hits a wall."""
    d, a = ACTIONS[action]
    front, back = state
    if d == 0:
        return state
    ang = np.arctan2(front[1] - back[1], front[0] - back[0]) + a
    dx, dy = int(d * np.rint(np.cos(ang))), int(d * np.rint(np.sin(ang)))
    if d > 0:
        new_back, new_front = front, (front[0] + dx, front[1] + dy)
    else:
        new_front, new_back = back, (back[0] + dx, back[1] + dy)
    if new_front in walls or new_back in walls:
        return None
    return (new_front, new_back)


def _cells(state):
    front, back = state
    cells = {front, back}
    if abs(front[0] - back[0]) + abs(front[1] - back[1]) > 1:  # diagonal body also sweeps the two corner cells
        cells |= {(front[0], back[1]), (back[0], front[1])}
    return cells


def plan_strategy(n_vehicles: int = 4, max_steps: int = 40, seed: int = 0) -> Dict[str, List[Dict[str, Tuple[int, int]]]]:
    """Prioritized space-time BFS over every priority order; ``seed`` indexes the feasible orders sorted by
    (total steps, longest plan, order), so seed 0 is the shortest joint strategy."""
    from itertools import permutations

    ranked = []
    for order in permutations(range(n_vehicles)):
        try:
            plans = _plan_with_order(list(order), max_steps)
        except RuntimeError:
            continue
        lens = [len(v) for v in plans.values()]
        ranked.append((sum(lens), max(lens), order, plans))
    if not ranked:
        raise RuntimeError("no collision-free joint strategy found")
    ranked.sort(key=lambda r: r[:3])
    return ranked[seed % len(ranked)][3]


def _plan_with_order(order, max_steps):
    walls = wall_cells()
    reserved: List[List[set]] = []  # reserved[t] = list of cell sets of already planned agents at step t
    plans = {}
    for idx in order:
        cfg = AGENT_CONFIGS[idx]
        start = (cfg["init_state"]["front"], cfg["init_state"]["back"])
        goal = (cfg["goal"]["front"], cfg["goal"]["back"])

        def free(state, t):
            if t >= len(reserved):
                return True
            c = _cells(state)
            return all(not (c & other) for other in reserved[t])

        queue = deque([(start, 0)])
        parent = {(start, 0): None}
        found = None
        while queue:
            state, t = queue.popleft()
            if state == goal:
                found = (state, t)
                break
            if t >= max_steps:
                continue
            for action in (2, 1, 3, 5, 4, 6, 0):
                nxt = move(state, action, walls)
                if nxt is None or (nxt, t + 1) in parent or not free(nxt, t + 1):
                    continue
                # swapping cells with a planned agent within one step counts as a collision too
                if t < len(reserved) and any(_cells(nxt) & other for other in reserved[t]):
                    continue
                parent[(nxt, t + 1)] = (state, t)
                queue.append((nxt, t + 1))
        if found is None:
            raise RuntimeError("no collision-free strategy found for vehicle_%d" % idx)
        path = []
        node = found
        while node is not None:
            path.append(node[0])
            node = parent[node]
        path.reverse()
        for t, st in enumerate(path):
            while len(reserved) <= t:
                reserved.append([])
            reserved[t].append(_cells(st))
        plans["vehicle_%d" % idx] = [{"front": st[0], "back": st[1]} for st in path]
    return {k: plans[k] for k in sorted(plans)}


def write_strategy(file_name: str, n_vehicles: int = 4, seed: int = 0, min_sets: int = 0):
    """Write ``<file_name>.pkl`` with the reference schema; optionally pad every agent with stops to ``min_sets``."""
    plans = plan_strategy(n_vehicles=n_vehicles, seed=seed)
    for agent, lst in plans.items():
        while len(lst) < min_sets:
            lst.append(dict(lst[-1]))
    with open(file_name + ".pkl", "wb") as f:
        pickle.dump(plans, f)
    return plans
