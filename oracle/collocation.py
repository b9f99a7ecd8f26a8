"""ORACLE (test infrastructure, never imported by the product path).

Collocation constants and the kinematic bicycle, restated from the reference.

* ``collocation_coefficients`` follows confrez/control/vehicle.py:54-97 (Lagrange basis on
  tau = [0, Radau(K)]; A[j,k] = L_j'(tau_k), B[j] = int_0^1 L_j, D[j] = L_j(1)).
  The reference gets the Radau points from CasADi (``ca.collocation_points(K, "radau")``,
  vehicle.py:66); CasADi is absent here, so they are recomputed from their definition
  (roots of P_K - P_{K-1} mapped to [0,1]) and pinned against SURVEY.md App. A.2.
* ``f_ct`` follows confrez/control/dynamic_model.py:5-27; ``f_rk4`` follows :30-58.
"""
import numpy as np
from numpy.polynomial import legendre


def radau_points(K: int) -> np.ndarray:
    c = np.zeros(K + 1)
    c[K], c[K - 1] = 1.0, -1.0
    return (np.sort(legendre.legroots(c)) + 1.0) / 2.0


def collocation_coefficients(K: int):
    tau_root = np.append(0, radau_points(K))
    A = np.zeros((K + 1, K + 1))
    D = np.zeros(K + 1)
    B = np.zeros(K + 1)
    for j in range(K + 1):
        p = np.poly1d([1])
        for k in range(K + 1):
            if k != j:
                p *= np.poly1d([1, -tau_root[k]]) / (tau_root[j] - tau_root[k])
        D[j] = p(1.0)
        pder = np.polyder(p)
        for k in range(K + 1):
            A[j, k] = pder(tau_root[k])
        B[j] = np.polyint(p)(1.0)
    return A, B, D


def f_ct(z, u, wb=2.5):
    """Continuous-time bicycle, state [x, y, psi, v, delta], input [a, w]."""
    x, y, psi, v, delta = z
    a, w = u
    return np.array([v * np.cos(psi), v * np.sin(psi), v / wb * np.tan(delta), a, w])


def f_rk4(z, u, dt, wb=2.5, M=4):
    """RK4 with M sub-steps of h = dt / M (dynamic_model.py:30-58)."""
    h = dt / M
    z = np.asarray(z, dtype=float)
    for _ in range(M):
        a1 = f_ct(z, u, wb)
        a2 = f_ct(z + h * a1 / 2, u, wb)
        a3 = f_ct(z + h * a2 / 2, u, wb)
        a4 = f_ct(z + h * a3, u, wb)
        z = z + h / 6 * (a1 + 2 * a2 + 2 * a3 + a4)
    return z
