"""Circle cover of the vehicle rectangle (reference: confrez/control/rect2circles.py:13-62) and the pair rows the circle variant of
the joint problem would impose (confrez/control/multi_vehicle_planner.py:153-179).

``num_circles`` discs of radius ``w / 2`` sit on the body's centre line, their centres spaced evenly from ``-cr`` to ``cf`` ahead of
the rear axle (``cr = -0.2``, ``cf = 2.45``, four discs for the default body).  Everything here is plain vectorised numpy on
``(x, y, psi)`` arrays: the functions serve as geometry helpers and as an independent clearance check of OBCA plans.  The NLP variant
itself (``solve_final_problem_circles``) is not built -- see ``MultiVehiclePlanner.solve_final_problem_circles``.
"""
import numpy as np

from conflict_rez_b200.vehicle_types import VehicleBody


def circle_offsets(vehicle_body: VehicleBody) -> np.ndarray:
    """Signed distances of the disc centres from the rear axle along the heading, shape (num_circles,)."""
    return np.linspace(-vehicle_body.cr, vehicle_body.cf, int(vehicle_body.num_circles))


def v2c_ca(x, y, psi, vehicle_body: VehicleBody):
    """Disc centres for poses of any shape: returns ``(xcs, ycs)`` with a trailing axis of length ``num_circles``
    (the numeric counterpart of the reference's symbolic ``v2c_ca``, rect2circles.py:13-37)."""
    x, y, psi = np.asarray(x, dtype=float), np.asarray(y, dtype=float), np.asarray(psi, dtype=float)
    o = circle_offsets(vehicle_body)
    return x[..., None] + np.cos(psi)[..., None] * o, y[..., None] + np.sin(psi)[..., None] * o


def v2c(state, vehicle_body: VehicleBody):
    """List of ``(xc, yc, radius)`` for one ``VehicleState`` (rect2circles.py:40-62)."""
    xcs, ycs = v2c_ca(state.x.x, state.x.y, state.e.psi, vehicle_body)
    r = vehicle_body.w / 2
    return [(float(a), float(b), r) for a, b in zip(xcs, ycs)]


def circle_pair_rows(xa, ya, psia, xb, yb, psib, vehicle_body: VehicleBody, d_buffer: float = 0.2) -> np.ndarray:
    """Values ``|c_a,j1 - c_b,j2|^2 - (w + d_buffer)^2`` of the circle variant's pair rows (multi_vehicle_planner.py:166-179;
    feasible when >= 0), shape ``poses.shape + (num_circles, num_circles)``."""
    ax, ay = v2c_ca(xa, ya, psia, vehicle_body)
    bx, by = v2c_ca(xb, yb, psib, vehicle_body)
    dx, dy = ax[..., :, None] - bx[..., None, :], ay[..., :, None] - by[..., None, :]
    return dx * dx + dy * dy - (vehicle_body.w + d_buffer) ** 2
