"""Region / obstacle containers mirroring ``confrez/obstacle_types.py``.

Reference: confrez/obstacle_types.py:10-25 (GeofenceRegion), :58-106
(BasePolytopeObstacle), :110-170 (RectangleObstacle).  No matplotlib.
"""
from dataclasses import dataclass, field

import numpy as np

from conflict_rez_b200.pytypes import PythonMsg


@dataclass
class GeofenceRegion:
    x_max: float = field(default=13 * 2.5)
    x_min: float = field(default=2.5)
    y_max: float = field(default=11 * 2.5)
    y_min: float = field(default=3 * 2.5)

    def xy(self):
        return np.array(
            [
                [self.x_max, self.y_max],
                [self.x_max, self.y_min],
                [self.x_min, self.y_min],
                [self.x_min, self.y_max],
                [self.x_max, self.y_max],
            ]
        )


@dataclass
class BaseObstacle(PythonMsg):
    xy: np.ndarray = field(default=None)


@dataclass
class BasePolytopeObstacle(BaseObstacle):
    """2-D convex obstacle kept both as vertices ``V`` and half-planes ``{x: A x <= b}``."""

    V: np.ndarray = field(default=None)
    A: np.ndarray = field(default=None)
    b: np.ndarray = field(default=None)

    def __setattr__(self, key, value):
        PythonMsg.__setattr__(self, key, value)
        self.__calc_V__()
        self.__calc_A_b__()

    def __calc_V__(self):  # pragma: no cover - abstract
        return

    def __calc_A_b__(self):  # pragma: no cover - abstract
        return


@dataclass
class RectangleObstacle(BasePolytopeObstacle):
    """Rectangle centred at (xc, yc), size w x h, rotated by psi (obstacle_types.py:110-170)."""

    xc: float = field(default=0)
    yc: float = field(default=0)
    w: float = field(default=0)
    h: float = field(default=0)
    psi: float = field(default=0)

    def __post_init__(self):
        self.__calc_V__()
        self.__calc_A_b__()

    def R(self):
        c, s = np.cos(self.psi), np.sin(self.psi)
        return np.array([[c, s], [-s, c]])

    def __calc_V__(self):
        hw, hh = self.w / 2, self.h / 2
        corners = np.array([[-hw, -hh], [-hw, hh], [hw, hh], [hw, -hh], [-hw, -hh]])
        xy = corners @ self.R() + np.array([[self.xc, self.yc]])
        object.__setattr__(self, "xy", xy)
        object.__setattr__(self, "V", xy[:-1, :])

    def __calc_A_b__(self):
        R = self.R()
        A = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]]) @ R
        centre = R @ np.array([self.xc, self.yc])  # centre expressed in the rectangle frame
        half = np.array([self.w / 2, self.h / 2])
        b = np.concatenate([centre + half, -centre + half])
        object.__setattr__(self, "A", A)
        object.__setattr__(self, "b", b)
