"""Angle helper kept from ``confrez/control/utils.py:28-29`` (plotting helpers are out of scope)."""
from math import pi


def pi_2_pi(angle):
    return (angle + pi) % (2 * pi) - pi
