"""Vehicle body and limits mirroring ``confrez/vehicle_types.py``.

Reference: confrez/vehicle_types.py:9-71 (VehicleBody: rear-axle referenced
rectangle, G = [[1,0],[0,1],[-1,0],[0,-1]], g = [lf, w/2, lr, w/2]) and :75-90
(VehicleConfig limits).
"""
from dataclasses import dataclass, field

import numpy as np

from conflict_rez_b200.pytypes import PythonMsg
from conflict_rez_b200.obstacle_types import BasePolytopeObstacle


@dataclass
class VehicleBody(BasePolytopeObstacle):
    hf: float = field(default=0.8)  # front overhang
    wb: float = field(default=2.5)  # wheelbase
    hr: float = field(default=0.6)  # rear overhang
    offset: float = field(default=0)
    lf: float = field(default=0)  # rear axle -> front bumper
    lr: float = field(default=0)  # rear axle -> rear bumper
    l: float = field(default=0)
    w: float = field(default=1.8)
    cr: float = field(default=0)
    cf: float = field(default=0)
    num_circles: int = field(default=3)

    def __post_init__(self):
        object.__setattr__(self, "offset", self.wb / 2)
        object.__setattr__(self, "lf", self.wb + self.hf)
        object.__setattr__(self, "lr", self.hr)
        object.__setattr__(self, "l", self.wb + self.hf + self.hr)
        object.__setattr__(self, "cf", 2.45)
        object.__setattr__(self, "cr", -0.2)
        object.__setattr__(self, "num_circles", 4)
        self.__calc_V__()
        self.__calc_A_b__()

    def __calc_V__(self):
        xy = np.array(
            [
                [self.lf, self.w / 2],
                [-self.lr, self.w / 2],
                [-self.lr, -self.w / 2],
                [self.lf, -self.w / 2],
                [self.lf, self.w / 2],
            ]
        )
        object.__setattr__(self, "xy", xy)
        object.__setattr__(self, "V", xy[:-1, :])

    def __calc_A_b__(self):
        object.__setattr__(self, "A", np.array([[1, 0], [0, 1], [-1, 0], [0, -1]]))
        object.__setattr__(self, "b", np.array([self.lf, self.w / 2, self.lr, self.w / 2]))


@dataclass
class VehicleConfig(PythonMsg):
    v_max: float = field(default=2.5)
    v_min: float = field(default=-2.5)
    a_max: float = field(default=1.5)
    a_min: float = field(default=-1.5)
    delta_max: float = field(default=0.85)
    delta_min: float = field(default=-0.85)
    w_delta_max: float = field(default=1)
    w_delta_min: float = field(default=-1)
