"""Cubic Bezier pose interpolation used for the ``spline_ws`` initial guess.

Reference: confrez/control/bezier.py:22-56 (BezierPlanner.interpolate): control
points [s, s + d*dir(s), e - d*dir(e), e] with d = |s-e|/offset, N samples of
t in [0,1) and heading from the first derivative.
"""
import numpy as np


class BezierPlanner(object):
    def __init__(self, offset: float):
        self.offset = offset

    def interpolate(self, start_state, end_state, N):
        s = np.array([start_state.x.x, start_state.x.y])
        e = np.array([end_state.x.x, end_state.x.y])
        syaw, eyaw = start_state.e.psi, end_state.e.psi
        dist = np.hypot(*(s - e)) / self.offset
        cp = np.array(
            [
                s,
                s + dist * np.array([np.cos(syaw), np.sin(syaw)]),
                e - dist * np.array([np.cos(eyaw), np.sin(eyaw)]),
                e,
            ]
        )
        t = np.linspace(0, 1, N, endpoint=False)[:, None]
        xy = (
            (1 - t) ** 3 * cp[0]
            + 3 * (1 - t) ** 2 * t * cp[1]
            + 3 * (1 - t) * t ** 2 * cp[2]
            + t ** 3 * cp[3]
        )
        d = 3 * ((1 - t) ** 2 * (cp[1] - cp[0]) + 2 * (1 - t) * t * (cp[2] - cp[1]) + t ** 2 * (cp[3] - cp[2]))
        yaw = np.arctan2(d[:, 1], d[:, 0])
        return np.column_stack([xy, yaw])
