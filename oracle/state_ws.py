"""ORACLE (test infrastructure, never imported by the product path).

Sparse restatement of the tube-following state warm start ``Vehicle.state_ws`` (confrez/control/vehicle.py:99-231):

    nodes k = 0 .. N*M (M = num_sets - 1), states z_k = (x, y, psi, v, delta), inputs u_k = (a, w) for k < N*M
    minimise    sum_k a_k^2 + w_k^2                                                   (:175-176)
    subject to  z_0 = init_state + init_offset (pose), v_0 = delta_0 = 0, a_0 = w_0 = 0   (:131-138)
                z_{k+1} = z_k + dt f(z_k, u_k)          forward Euler, kinematic bicycle      (:169-173)
                region / v / delta bounds at k < N*M, a / w bounds only if bounded_input      (:140-167)
                A_back (x, y) <= b - shrink,  A_front (x + wb cos psi, y + wb sin psi) <= b - shrink   at k = N i, i = 1..M   (:178-192)
                psi_{N M} = final_heading (if given)                                          (:194-195)

Inequalities carry slacks like IPOPT's own formulation (g(x) - s = 0, s >= 0).  Derivatives are written out analytically here
(the CUDA kernel gets them from Taylor jets); tests/test_state_ws.py checks them against finite differences.  Solved with
oracle/ipm.py like the other oracle problems.
"""
import numpy as np
import scipy.sparse as sp

from oracle.nlp import _Alloc

INF = np.inf


class EulerWsNLP:
    """prob: dict(N, dt, wb, tube_A (S,2,4,2), tube_b (S,2,4) [b already reduced by shrink_tube; set 0 unused], init (3,),
    heading (float or None), region (4,), limits (8,), bounded_input (bool))."""

    def __init__(self, prob):
        self.p = prob
        self.S = prob["tube_A"].shape[0]
        self.M = self.S - 1
        self.NM = prob["N"] * self.M
        K = self.NM + 1
        va = _Alloc()
        self.iz = va.take(K, 5)
        self.iu = va.take(self.NM, 2)
        self.its = va.take(self.M, 8)
        self.n = va.n
        xL, xU = np.full(self.n, -INF), np.full(self.n, INF)
        rg, lm = prob["region"], prob["limits"]
        lo, hi = [rg[0], rg[2], -INF, lm[0], lm[2]], [rg[1], rg[3], INF, lm[1], lm[3]]
        xL[self.iz[:-1]], xU[self.iz[:-1]] = lo, hi  # the last node carries no bounds (range(N * M), vehicle.py:140)
        if prob.get("bounded_input"):
            xL[self.iu], xU[self.iu] = [lm[4], lm[6]], [lm[5], lm[7]]
        xL[self.its] = 0.0
        self.xL, self.xU = xL, xU
        ra = _Alloc()
        self.r_init = ra.take(7)
        self.r_dyn = ra.take(self.NM, 5)
        self.r_tube = ra.take(self.M, 8)
        self.r_head = ra.take(1) if prob.get("heading") is not None else np.zeros(0, dtype=int)
        self.m = ra.n
        self.knode = prob["N"] * np.arange(1, self.S)  # nodes that carry a tube set

    # -- pieces
    def _f(self, z, u):
        wb = self.p["wb"]
        return np.stack([z[:, 3] * np.cos(z[:, 2]), z[:, 3] * np.sin(z[:, 2]), z[:, 3] / wb * np.tan(z[:, 4]), u[:, 0], u[:, 1]], axis=1)

    def _tube_rows(self):
        """(M, 8, 3): row (ax, ay, b) of set i = 1..M; rows 0..3 act on the rear axle, 4..7 on the front axle."""
        A, b = self.p["tube_A"][1:], self.p["tube_b"][1:]
        return np.concatenate([A.reshape(self.M, 8, 2), b.reshape(self.M, 8, 1)], axis=2)

    def f(self, x):
        u = x[self.iu]
        return float((u ** 2).sum())

    def grad_f(self, x):
        g = np.zeros(self.n)
        g[self.iu] = 2.0 * x[self.iu]
        return g

    def c(self, x):
        out = np.zeros(self.m)
        z, u = x[self.iz], x[self.iu]
        out[self.r_init[:3]] = z[0, :3] - self.p["init"]
        out[self.r_init[3:5]] = z[0, 3:5]
        out[self.r_init[5:]] = u[0]
        out[self.r_dyn] = z[1:] - z[:-1] - self.p["dt"] * self._f(z[:-1], u)
        t = self._tube_rows()
        zk = z[self.knode]
        wb = self.p["wb"]
        px = np.concatenate([np.repeat(zk[:, 0:1], 4, 1), np.repeat(zk[:, 0:1] + wb * np.cos(zk[:, 2:3]), 4, 1)], axis=1)
        py = np.concatenate([np.repeat(zk[:, 1:2], 4, 1), np.repeat(zk[:, 1:2] + wb * np.sin(zk[:, 2:3]), 4, 1)], axis=1)
        out[self.r_tube] = t[:, :, 2] - t[:, :, 0] * px - t[:, :, 1] * py - x[self.its]
        if len(self.r_head):
            out[self.r_head] = z[-1, 2] - self.p["heading"]
        return out

    def jac(self, x):
        rr, cc, vv = [], [], []

        def put(r, c, v):
            r, c = np.broadcast_arrays(np.asarray(r), np.asarray(c))
            rr.append(r.ravel()), cc.append(c.ravel()), vv.append(np.broadcast_to(np.asarray(v, dtype=float), r.shape).ravel())

        z, u = x[self.iz], x[self.iu]
        dt, wb = self.p["dt"], self.p["wb"]
        put(self.r_init[:5], self.iz[0], 1.0)
        put(self.r_init[5:], self.iu[0], 1.0)
        for r in range(5):
            put(self.r_dyn[:, r], self.iz[1:, r], 1.0)
            put(self.r_dyn[:, r], self.iz[:-1, r], -1.0)
        zz = z[:-1]
        cs, sn, tn = np.cos(zz[:, 2]), np.sin(zz[:, 2]), np.tan(zz[:, 4])
        put(self.r_dyn[:, 0], self.iz[:-1, 2], dt * zz[:, 3] * sn)
        put(self.r_dyn[:, 0], self.iz[:-1, 3], -dt * cs)
        put(self.r_dyn[:, 1], self.iz[:-1, 2], -dt * zz[:, 3] * cs)
        put(self.r_dyn[:, 1], self.iz[:-1, 3], -dt * sn)
        put(self.r_dyn[:, 2], self.iz[:-1, 3], -dt * tn / wb)
        put(self.r_dyn[:, 2], self.iz[:-1, 4], -dt * zz[:, 3] * (1 + tn * tn) / wb)
        put(self.r_dyn[:, 3], self.iu[:, 0], -dt)
        put(self.r_dyn[:, 4], self.iu[:, 1], -dt)
        t = self._tube_rows()
        zk = z[self.knode]
        for r in range(8):
            put(self.r_tube[:, r], self.iz[self.knode, 0], -t[:, r, 0])
            put(self.r_tube[:, r], self.iz[self.knode, 1], -t[:, r, 1])
            if r >= 4:
                put(self.r_tube[:, r], self.iz[self.knode, 2], -wb * (-t[:, r, 0] * np.sin(zk[:, 2]) + t[:, r, 1] * np.cos(zk[:, 2])))
            put(self.r_tube[:, r], self.its[:, r], -1.0)
        if len(self.r_head):
            put(self.r_head, self.iz[-1, 2], 1.0)
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(self.m, self.n))

    def hess(self, x, y, clip=True):
        rr, cc, vv = [], [], []

        def put(r, c, v):
            rr.append(np.asarray(r).ravel()), cc.append(np.asarray(c).ravel()), vv.append(np.asarray(v, dtype=float).ravel())

        z = x[self.iz]
        dt, wb = self.p["dt"], self.p["wb"]
        put(self.iu, self.iu, np.full(self.iu.shape, 2.0))
        zz, yd = z[:-1], y[self.r_dyn]
        cs, sn, tn = np.cos(zz[:, 2]), np.sin(zz[:, 2]), np.tan(zz[:, 4])
        sec2 = 1 + tn * tn
        # rows are -dt f_r: second derivatives of f (SURVEY.md A.4 checklist)
        hpp = -dt * (yd[:, 0] * (-zz[:, 3] * cs) + yd[:, 1] * (-zz[:, 3] * sn))
        hpv = -dt * (yd[:, 0] * (-sn) + yd[:, 1] * cs)
        hvd = -dt * yd[:, 2] * sec2 / wb
        hdd = -dt * yd[:, 2] * 2 * zz[:, 3] * sec2 * tn / wb
        i2, i3, i4 = self.iz[:-1, 2], self.iz[:-1, 3], self.iz[:-1, 4]
        put(i2, i2, hpp), put(i2, i3, hpv), put(i3, i2, hpv), put(i3, i4, hvd), put(i4, i3, hvd), put(i4, i4, hdd)
        t = self._tube_rows()
        zk, yt = z[self.knode], y[self.r_tube]
        hk = np.zeros(self.M)
        for r in range(4, 8):
            hk += yt[:, r] * wb * (t[:, r, 0] * np.cos(zk[:, 2]) + t[:, r, 1] * np.sin(zk[:, 2]))
        put(self.iz[self.knode, 2], self.iz[self.knode, 2], hk)
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(self.n, self.n))

    def init_slacks(self, x):
        x = x.copy()
        x[self.its] = 0.0
        x[self.its] = self.c(x)[self.r_tube]
        return x

    def pack(self, z, u):
        x = np.zeros(self.n)
        x[self.iz], x[self.iu] = z, u
        return x
