"""Message dataclasses mirroring ``confrez/pytypes.py`` (field names and freeze semantics kept).

Reference: confrez/pytypes.py:13-90 (PythonMsg), :153-217 (small structs),
:355-451 (VehicleState), :456-517 (VehiclePrediction).  Unlike the reference this
module has no import-time matplotlib/pdb dependency.
Only the fields the OBCA planners read or write are kept; unused racing-specific
fields (parametric pose, covariances, quaternion) are dropped.
"""
from dataclasses import dataclass, field, fields
import copy
from typing import Any

import numpy as np


@dataclass
class PythonMsg:
    """Dataclass base whose instances refuse creation of unknown attributes (pytypes.py:24-37)."""

    def __setattr__(self, key, value):
        if not hasattr(self, key):
            raise TypeError('Cannot add new field "%s" to frozen class %s' % (key, self))
        object.__setattr__(self, key, value)

    def print(self, depth=0, name=None):
        pad = "  " * depth
        out = pad + ((name + " (" + type(self).__name__ + "):\n") if name else type(self).__name__ + ":\n")
        for key in vars(self):
            val = getattr(self, key)
            if isinstance(val, PythonMsg):
                out += val.print(depth=depth + 1, name=key)
            else:
                out += "  " * (depth + 1) + "%s=%s\n" % (key, val)
        if depth == 0:
            print(out)
            return None
        return out

    def copy(self):
        return copy.deepcopy(self)


@dataclass
class Position(PythonMsg):
    x: float = field(default=0)
    y: float = field(default=0)
    z: float = field(default=0)


@dataclass
class VehicleActuation(PythonMsg):
    t: float = field(default=0)
    u_a: float = field(default=0)
    u_steer: float = field(default=0)
    u_steer_dot: float = field(default=0)


@dataclass
class BodyLinearVelocity(PythonMsg):
    v_long: float = field(default=0)
    v_tran: float = field(default=0)
    v_n: float = field(default=0)
    v: float = field(default=0)  # speed of the bicycle model

    def mag(self):
        return np.sqrt(self.v_long ** 2 + self.v_tran ** 2 + self.v_n ** 2)


@dataclass
class BodyAngularVelocity(PythonMsg):
    w_phi: float = field(default=0)
    w_theta: float = field(default=0)
    w_psi: float = field(default=0)


@dataclass
class BodyLinearAcceleration(PythonMsg):
    a_long: float = field(default=0)
    a_tran: float = field(default=0)
    a_n: float = field(default=0)


@dataclass
class OrientationEuler(PythonMsg):
    phi: float = field(default=0)
    theta: float = field(default=0)
    psi: float = field(default=0)


@dataclass
class VehicleState(PythonMsg):
    """Vehicle state: ``x`` position, ``e.psi`` heading, ``v.v`` speed, ``u`` actuation (pytypes.py:355-404)."""

    vehicle_id: int = field(default=1)
    t: float = field(default=None)
    x: Position = field(default=None)
    v: BodyLinearVelocity = field(default=None)
    w: BodyAngularVelocity = field(default=None)
    a: BodyLinearAcceleration = field(default=None)
    e: OrientationEuler = field(default=None)
    u: VehicleActuation = field(default=None)

    def __post_init__(self):
        if self.x is None:
            self.x = Position()
        if self.u is None:
            self.u = VehicleActuation()
        if self.v is None:
            self.v = BodyLinearVelocity()
        if self.w is None:
            self.w = BodyAngularVelocity()
        if self.a is None:
            self.a = BodyLinearAcceleration()
        if self.e is None:
            self.e = OrientationEuler()

    def get_R(self, reverse=False):
        psi = -self.e.psi if reverse else self.e.psi
        return np.array([[np.cos(psi), -np.sin(psi), 0], [np.sin(psi), np.cos(psi), 0], [0, 0, 0]])


@dataclass
class VehiclePrediction(PythonMsg):
    """Trajectory container: flat time-ordered arrays (pytypes.py:456-517).

    ``l``/``m`` hold the OBCA obstacle duals; their layout depends on the stage
    that produced them.
    """

    t: Any = field(default=None)
    dt: float = field(default=None)
    x: Any = field(default=None)
    y: Any = field(default=None)
    v: Any = field(default=None)
    l: Any = field(default=None)
    m: Any = field(default=None)
    psi: Any = field(default=None)
    psidot: Any = field(default=None)
    u_a: Any = field(default=None)
    u_steer: Any = field(default=None)
    u_steer_dot: Any = field(default=None)
