"""Minimal 2-D convex polytope replacing the external ``pytope.Polytope``.

The reference builds obstacles and tube sets with ``pytope.Polytope(V)`` and
reads ``.A`` (m x 2), ``.b`` (m x 1 column), ``.V`` and ``P + offset``
(confrez/control/compute_sets.py:7,34-39,136; vehicle.py:182-184,529-541).
pytope is not vendored with the reference; this class offers the same surface
with **unit outward normals** listed counter-clockwise starting from the edge
leaving the lowest-then-leftmost vertex.  Row order/scale only changes the
obstacle duals, never a primal quantity.
"""
import numpy as np


def _convex_hull_ccw(points: np.ndarray) -> np.ndarray:
    pts = sorted(set(map(tuple, np.asarray(points, dtype=float))))
    if len(pts) < 3:
        raise ValueError("a 2-D polytope needs at least 3 distinct vertices")

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1])


class Polytope:
    """Convex polygon ``{x : A x <= b}`` built from its vertices."""

    def __init__(self, V):
        self.V = _convex_hull_ccw(np.asarray(V, dtype=float))
        nv = len(self.V)
        A = np.zeros((nv, 2))
        b = np.zeros((nv, 1))
        for i in range(nv):
            p, q = self.V[i], self.V[(i + 1) % nv]
            e = q - p
            nrm = np.array([e[1], -e[0]])  # outward normal of a CCW polygon
            nrm = nrm / np.linalg.norm(nrm)
            A[i] = nrm
            b[i, 0] = nrm @ p
        self.A = A
        self.b = b

    def __add__(self, offset):
        return Polytope(self.V + np.asarray(offset, dtype=float).reshape(1, 2))

    __radd__ = __add__

    def contains(self, x, tol=1e-9):
        return bool(np.all(self.A @ np.asarray(x, dtype=float).reshape(2) <= self.b[:, 0] + tol))

    def __repr__(self):
        return "Polytope(V=%s)" % (np.array2string(self.V, precision=3),)
