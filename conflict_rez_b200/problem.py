"""Problem descriptors filled by the planners *instead of* building a CasADi graph.

``CollocationProblem`` carries exactly the data that
``Vehicle.setup_single_final_problem`` (confrez/control/vehicle.py:360-640) and
``MultiVehiclePlanner.solve_final_problem_obca`` (multi_vehicle_planner.py:343-480) bake into
their ``ca.Opti`` object: collocation shape, obstacles (H-representation), body, limits, tube
sets, initial poses, final headings, ``dmin``/``shrink_tube``.  ``CollocationGuess`` is the
``opti.set_initial`` data (state/input guess, obstacle duals, pair duals, ``dt0``).

All arrays are FP64, C-contiguous, vehicle-major then node-major; ragged vehicles are padded to
the longest horizon (``n_sets`` tells the real length).  A leading batch axis ``B`` may be present
on per-instance fields (``init_pose`` and every guess array).
"""
from dataclasses import dataclass, field
from itertools import combinations
from typing import List, Optional

import numpy as np

H_FACES = 4  # every obstacle / tube set / body is a 4-face polytope in the reference scenarios


@dataclass
class CollocationProblem:
    n_sets: np.ndarray  # (V,) int: strategy sets per vehicle (S_a); N_a = n_per_set * (S_a - 1)
    obs_A: np.ndarray  # (O, 4, 2)
    obs_b: np.ndarray  # (O, 4)
    tube_A: np.ndarray  # (V, Smax, 2[back,front], 4, 2)
    tube_b: np.ndarray  # (V, Smax, 2, 4)  raw b (shrink_tube not yet subtracted)
    init_pose: np.ndarray  # ([B,] V, 3): x, y, psi *including* init_offset
    final_heading: np.ndarray  # (V,) nan = unconstrained
    body_G: np.ndarray = field(default_factory=lambda: np.array([[1.0, 0], [0, 1], [-1, 0], [0, -1]]))
    body_g: np.ndarray = field(default_factory=lambda: np.array([3.3, 0.9, 0.6, 0.9]))
    wb: float = 2.5
    region: np.ndarray = field(default_factory=lambda: np.array([2.5, 32.5, 7.5, 27.5]))  # xmin xmax ymin ymax
    limits: np.ndarray = field(default_factory=lambda: np.array([-2.5, 2.5, -0.85, 0.85, -1.5, 1.5, -1.0, 1.0]))  # v, delta, a, w (min,max)
    K: int = 5
    n_per_set: int = 5
    dmin: float = 0.05
    shrink_tube: float = 0.5

    def __post_init__(self):
        self.n_sets = np.asarray(self.n_sets, dtype=np.int64)
        for name in ("obs_A", "obs_b", "tube_A", "tube_b", "init_pose", "final_heading", "body_G", "body_g", "region", "limits"):
            setattr(self, name, np.ascontiguousarray(np.asarray(getattr(self, name), dtype=np.float64)))
        assert self.obs_A.shape[1:] == (H_FACES, 2) and self.obs_b.shape == self.obs_A.shape[:2]
        assert self.tube_A.shape[0] == self.V and self.tube_A.shape[2:] == (2, H_FACES, 2)

    @property
    def V(self) -> int:
        return len(self.n_sets)

    @property
    def O(self) -> int:
        return self.obs_A.shape[0]

    @property
    def N(self) -> np.ndarray:
        """Collocation intervals per vehicle (vehicle.py:381)."""
        return self.n_per_set * (self.n_sets - 1)

    @property
    def nodes(self) -> np.ndarray:
        return self.N * (self.K + 1)

    @property
    def pairs(self) -> List[tuple]:
        return list(combinations(range(self.V), 2))

    @property
    def batch(self) -> Optional[int]:
        return self.init_pose.shape[0] if self.init_pose.ndim == 3 else None

    def instance(self, b: int) -> "CollocationProblem":
        """Single-instance view of a batched problem."""
        if self.batch is None:
            return self
        import copy

        p = copy.copy(self)
        p.init_pose = self.init_pose[b]
        return p


@dataclass
class CollocationGuess:
    """Warm start (``opti.set_initial`` data).  Shapes with Mmax = max nodes, P = number of vehicle pairs."""

    z: np.ndarray  # ([B,] V, Mmax, 7): x, y, psi, v, delta, a, w at every collocation node
    lam: np.ndarray  # ([B,] V, Mmax, O, 4) obstacle duals lambda
    mu: np.ndarray  # ([B,] V, Mmax, O, 4) obstacle duals mu
    dt: np.ndarray  # ([B,]) initial interval length
    pair_lam: Optional[np.ndarray] = None  # ([B,] P, Mmax, 4)  lambda_ij
    pair_mu: Optional[np.ndarray] = None  # ([B,] P, Mmax, 4)  lambda_ji
    pair_s: Optional[np.ndarray] = None  # ([B,] P, Mmax, 2)

    def instance(self, b: int) -> "CollocationGuess":
        if np.ndim(self.dt) == 0:
            return self
        pick = lambda a: None if a is None else a[b]
        return CollocationGuess(self.z[b], self.lam[b], self.mu[b], self.dt[b], pick(self.pair_lam), pick(self.pair_mu), pick(self.pair_s))
