"""Host-side warm starts for the collocation OBCA solve (reference: warm-start stages of the planners).

* ``kinematic_guess``      pose samples -> dynamically plausible (x, y, psi, v, delta, a, w); plays the role of
                           ``Vehicle.state_ws`` (confrez/control/vehicle.py:99-231) for the synthetic benchmarks:
                           the reference solves an Euler-discretised NLP there, here the Bezier pose guess of
                           ``interp_along_sets`` (compute_sets.py:167-240) is differentiated instead.
* ``dual_ws_rect``         closed form of ``Vehicle.dual_ws`` (vehicle.py:233-296) for 4-face rectangles: the problem
                           is the dual of the rectangle-rectangle distance, so lambda = max(0, A w), mu = max(0, -G R'w)
                           with w the unit vector between the closest points.
* ``joint_dual_ws_rect``   closed form of ``MultiVehiclePlanner.joint_dual_ws`` (multi_vehicle_planner.py:208-341).
* ``interp_ws_for_collocation``  linear resampling onto the Radau grid (vehicle.py:298-358).
"""
import numpy as np


def _rect_vertices(A, b):
    """Vertices (4,2) of {x: A x <= b} with 4 faces listed around the boundary (adjacent faces intersect)."""
    V = np.zeros(A.shape[:-2] + (4, 2))
    for i in range(4):
        j = (i + 1) % 4
        M = np.stack([A[..., i, :], A[..., j, :]], axis=-2)
        rhs = np.stack([b[..., i], b[..., j]], axis=-1)
        V[..., i, :] = np.linalg.solve(M, rhs[..., None])[..., 0]
    return V


def _point_segment(p, a, b):
    """Closest point on segment ab to p (broadcast over leading axes)."""
    ab = b - a
    t = np.clip(np.sum((p - a) * ab, -1) / np.maximum(np.sum(ab * ab, -1), 1e-300), 0.0, 1.0)
    return a + t[..., None] * ab


def closest_points_convex(P, Q):
    """Closest points between two non-overlapping convex quadrilaterals given as vertex loops (...,4,2).

    Returns (p on P, q on Q, distance); overlapping shapes are handled by ``_sat_direction``.
    """
    best_d = np.full(P.shape[:-2], np.inf)
    best_p = np.zeros(P.shape[:-2] + (2,))
    best_q = np.zeros_like(best_p)
    for i in range(4):
        for j in range(4):
            # vertex i of P against edge j of Q
            q = _point_segment(P[..., i, :], Q[..., j, :], Q[..., (j + 1) % 4, :])
            d = np.linalg.norm(P[..., i, :] - q, axis=-1)
            upd = d < best_d
            best_d = np.where(upd, d, best_d)
            best_p = np.where(upd[..., None], P[..., i, :], best_p)
            best_q = np.where(upd[..., None], q, best_q)
            # vertex i of Q against edge j of P
            p = _point_segment(Q[..., i, :], P[..., j, :], P[..., (j + 1) % 4, :])
            d = np.linalg.norm(Q[..., i, :] - p, axis=-1)
            upd = d < best_d
            best_d = np.where(upd, d, best_d)
            best_p = np.where(upd[..., None], p, best_p)
            best_q = np.where(upd[..., None], Q[..., i, :], best_q)
    return best_p, best_q, best_d


def _body_vertices(x, y, psi, G, g):
    Vb = _rect_vertices(G, g)  # body frame
    c, s = np.cos(psi), np.sin(psi)
    R = np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)  # (...,2,2)
    return np.einsum("...ij,kj->...ki", R, Vb) + np.stack([x, y], -1)[..., None, :]


def _unit(p, q, sat):
    """Unit vector q -> p for separated shapes, the least-penetration axis for overlapping ones."""
    sat_dir, sat_sep = sat
    w = p - q
    n = np.linalg.norm(w, axis=-1, keepdims=True)
    use_sat = (sat_sep[..., None] <= 1e-9) | (n <= 1e-9)
    return np.where(use_sat, sat_dir, w / np.maximum(n, 1e-300))


def _sat_direction(P, NP_, Q, NQ):
    """Least-penetration axis between overlapping convex quads: unit vector pointing from Q towards P.

    P, Q: vertex loops (...,4,2); NP_, NQ: outward unit face normals (...,4,2).  Candidate axes are the face
    normals of both shapes; the axis with the largest (least negative) separation is returned.
    """
    # faces of Q: separation = min_v nQ.(v_P) - max_u nQ.(u_Q)
    sepQ = np.einsum("...fc,...vc->...fv", NQ, P).min(-1) - np.einsum("...fc,...vc->...fv", NQ, Q).max(-1)
    sepP = np.einsum("...fc,...vc->...fv", NP_, Q).min(-1) - np.einsum("...fc,...vc->...fv", NP_, P).max(-1)
    iq, ip = sepQ.argmax(-1), sepP.argmax(-1)
    bestQ = np.take_along_axis(sepQ, iq[..., None], -1)[..., 0]
    bestP = np.take_along_axis(sepP, ip[..., None], -1)[..., 0]
    dq = np.take_along_axis(NQ, iq[..., None, None].repeat(2, -1), -2)[..., 0, :]
    dp = -np.take_along_axis(NP_, ip[..., None, None].repeat(2, -1), -2)[..., 0, :]
    return np.where((bestQ >= bestP)[..., None], dq, dp), np.maximum(bestQ, bestP)


def dual_ws_rect(x, y, psi, obs_A, obs_b, G, g):
    """Obstacle duals for poses (x,y,psi) of shape (...,): returns lam (..., O, 4), mu (..., O, 4)."""
    body = _body_vertices(x, y, psi, G, g)  # (...,4,2)
    O = obs_A.shape[0]
    lam = np.zeros(np.shape(x) + (O, 4))
    mu = np.zeros_like(lam)
    c, s = np.cos(psi), np.sin(psi)
    for j in range(O):
        Vo = np.broadcast_to(_rect_vertices(obs_A[j], obs_b[j]), body.shape)
        pb, po, _ = closest_points_convex(body, Vo)
        Rm = np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)
        nbody = np.einsum("...ij,kj->...ki", Rm, G)
        w = _unit(pb, po, _sat_direction(body, nbody, Vo, np.broadcast_to(obs_A[j], body.shape)))  # from obstacle towards body
        lam[..., j, :] = np.maximum(0.0, w @ obs_A[j].T)
        wb_ = np.stack([c * w[..., 0] + s * w[..., 1], -s * w[..., 0] + c * w[..., 1]], -1)  # R' w
        mu[..., j, :] = np.maximum(0.0, -(wb_ @ G.T))
    return lam, mu


def joint_dual_ws_rect(xa, ya, pa, xb, yb, pb, G, g):
    """Pair duals: lam (...,4), mu (...,4), s (...,2) with A_a'lam + s = 0, A_b'mu - s = 0, |s| = 1."""
    Ba = _body_vertices(xa, ya, pa, G, g)
    Bb = _body_vertices(xb, yb, pb, G, g)
    p, q, _ = closest_points_convex(Ba, Bb)

    def normals(psi):
        c, sn = np.cos(psi), np.sin(psi)
        Rm = np.stack([np.stack([c, -sn], -1), np.stack([sn, c], -1)], -2)
        return np.einsum("...ij,kj->...ki", Rm, G)

    s = _unit(p, q, _sat_direction(Ba, normals(pa), Bb, normals(pb)))  # from b towards a

    def to_body(psi, w):
        c, sn = np.cos(psi), np.sin(psi)
        return np.stack([c * w[..., 0] + sn * w[..., 1], -sn * w[..., 0] + c * w[..., 1]], -1)

    lam = np.maximum(0.0, to_body(pa, -s) @ G.T)
    mu = np.maximum(0.0, to_body(pb, s) @ G.T)
    return lam, mu, s


def kinematic_guess(path, dt, wb, limits):
    """(..., T, 3) pose samples spaced ``dt`` apart -> dict of arrays x,y,psi,v,delta,a,w (..., T) and t (T,)."""
    path = np.asarray(path, dtype=float)
    x, y, psi = path[..., 0], path[..., 1], np.unwrap(path[..., 2], axis=-1)
    T = x.shape[-1]
    dx, dy = np.gradient(x, dt, axis=-1), np.gradient(y, dt, axis=-1)
    v = dx * np.cos(psi) + dy * np.sin(psi)
    v = np.clip(v, 0.9 * limits[0], 0.9 * limits[1])
    v[..., 0] = v[..., -1] = 0.0
    # Steering is left at zero: differentiating the Bezier heading gives a steering guess that saturates around every
    # cusp and stop move (|v| ~ 0), which starts the interior-point iteration on the steering bounds and made ~5 % of
    # the randomized instances jam there.  delta = w = 0 starts strictly inside the box (the reference's state_ws NLP
    # also starts from zero steering, vehicle.py:197-205).
    delta = np.zeros_like(v)
    w = np.zeros_like(v)
    a = np.clip(np.gradient(v, dt, axis=-1), 0.9 * limits[4], 0.9 * limits[5])
    a[..., 0] = a[..., -1] = 0.0
    return {"t": dt * np.arange(T), "x": x, "y": y, "psi": psi, "v": v, "delta": delta, "a": a, "w": w}


def radau_nodes(K=5):
    from numpy.polynomial import legendre

    c = np.zeros(K + 1)
    c[K], c[K - 1] = 1.0, -1.0
    return np.append(0.0, (np.sort(legendre.legroots(c)) + 1.0) / 2.0)


def collocation_coefficients(K: int = 5):
    """Lagrange-basis collocation matrices on tau = [0, Radau(K)], computed the way the reference computes them
    (numpy poly1d arithmetic, confrez/control/vehicle.py:54-97): A[j,k] = L_j'(tau_k), B[j] = int_0^1 L_j, D[j] = L_j(1).
    The solver receives A and B through ObcaStatic, so host, oracle and kernels use bit-identical constants."""
    tau = radau_nodes(K)
    A, B, D = np.zeros((K + 1, K + 1)), np.zeros(K + 1), np.zeros(K + 1)
    for j in range(K + 1):
        p = np.poly1d([1.0])
        for k in range(K + 1):
            if k != j:
                p *= np.poly1d([1.0, -tau[k]]) / (tau[j] - tau[k])
        D[j] = p(1.0)
        A[j] = np.polyder(p)(tau)
        B[j] = np.polyint(p)(1.0)
    return A, B, D


def interp_ws_for_collocation(t, signals, N, K=5):
    """Linear interpolation of every signal (T,...) onto t_interp = (i + tau_k)/N * t[-1] (vehicle.py:321-331)."""
    tau = radau_nodes(K)
    t_interp = (np.arange(N)[:, None] + tau[None, :]).ravel() / N * t[-1]
    out = {}
    for name, sig in signals.items():
        sig = np.asarray(sig, dtype=float)
        flat = sig.reshape(len(t), -1)
        res = np.stack([np.interp(t_interp, t, flat[:, c]) for c in range(flat.shape[1])], axis=1)
        out[name] = res.reshape((len(t_interp),) + sig.shape[1:])
    return t_interp, out


def resample_for_collocation(t, sig, N, K=5):
    """Batched form of ``interp_ws_for_collocation`` for signals (..., T) that share the time grid ``t`` (T,):
    linear interpolation onto t_interp = (i + tau_k)/N * t[-1]; returns (..., N*(K+1))."""
    tau = radau_nodes(K)
    t_interp = (np.arange(N)[:, None] + tau[None, :]).ravel() / N * t[-1]
    idx = np.clip(np.searchsorted(t, t_interp, side="right") - 1, 0, len(t) - 2)
    wgt = (t_interp - t[idx]) / (t[idx + 1] - t[idx])
    sig = np.asarray(sig, dtype=float)
    return sig[..., idx] + (sig[..., idx + 1] - sig[..., idx]) * wgt
