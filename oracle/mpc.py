"""ORACLE (test infrastructure, never imported by the product path).

Sparse restatement of the distributed-MPC NLP of ``VehicleFollower.setup_controller``
(confrez/control/vehicle_follower.py:146-368): horizon N nodes, RK4 x 4 dynamics
(confrez/control/dynamic_model.py:30-58), obstacle OBCA triples at every node (:275-290), one pair block per other
vehicle with the other's predicted pose as a *parameter* (:322-352), tracking cost (:263-272).

    vars   z_i (5), u_i (2), lam/mu (O x 4 each), sd, el per node; per other: lam_ij, lam_ji (4), s (2), sd, sn, el
    rows   z_0 = current state (5); z_{i+1} - F(z_i, u_i) = 0 (i < N-1); obstacle rows (4); pair rows (6)

The dynamics derivatives are independent of the CUDA implementation (which uses second-order forward-mode jets): the
Jacobian is the analytic chain rule through the 16 stage evaluations, the multiplier-contracted Hessian is its
complex-step derivative.  Same slack / elastic conventions as oracle/nlp.py.
"""
import numpy as np
import scipy.sparse as sp

from oracle import blocks
from oracle.nlp import _Alloc

INF = np.inf


def _f(z, u, wb):
    """Continuous dynamics on arrays (5, n), (2, n); works for complex input."""
    return np.stack([z[3] * np.cos(z[2]), z[3] * np.sin(z[2]), z[3] / wb * np.tan(z[4]), u[0] + 0 * z[0], u[1] + 0 * z[0]])


def _df(z, u, wb):
    """df/d(z,u): (5, 7, n)."""
    n = z.shape[1]
    J = np.zeros((5, 7, n), dtype=z.dtype)
    c, s, t = np.cos(z[2]), np.sin(z[2]), np.tan(z[4])
    J[0, 2], J[0, 3] = -z[3] * s, c
    J[1, 2], J[1, 3] = z[3] * c, s
    J[2, 3], J[2, 4] = t / wb, z[3] * (1 + t * t) / wb
    J[3, 5] = 1.0
    J[4, 6] = 1.0
    return J


def rk4(z, u, dt, wb, M=4):
    """F(z,u) and dF/d(z,u) for arrays (5,n),(2,n) -> (5,n), (5,7,n): analytic chain rule through the stages."""
    h = dt / M
    n = z.shape[1]
    dz = np.zeros((5, 7, n), dtype=z.dtype)  # d z_k / d (z0, u)
    dz[np.arange(5), np.arange(5)] = 1.0
    du = np.zeros((2, 7, n), dtype=z.dtype)
    du[0, 5] = du[1, 6] = 1.0

    def stage(zz, dzz):
        J = _df(zz, u, wb)  # (5,7,n)
        d = np.einsum("ijn,jkn->ikn", J[:, :5], dzz) + np.einsum("ijn,jkn->ikn", J[:, 5:], du)
        return _f(zz, u, wb), d

    for _ in range(M):
        a1, d1 = stage(z, dz)
        a2, d2 = stage(z + h * a1 / 2, dz + h * d1 / 2)
        a3, d3 = stage(z + h * a2 / 2, dz + h * d2 / 2)
        a4, d4 = stage(z + h * a3, dz + h * d3)
        z = z + h / 6 * (a1 + 2 * a2 + 2 * a3 + a4)
        dz = dz + h / 6 * (d1 + 2 * d2 + 2 * d3 + d4)
    return z, dz


def rk4_hess(z, u, y, dt, wb):
    """sum_r y_r d2F_r/d(z,u)2 by complex-step differentiation of the analytic Jacobian: (7,7,n)."""
    n = z.shape[1]
    H = np.zeros((7, 7, n))
    hstep = 1e-30
    zu = np.concatenate([z, u]).astype(complex)
    for k in range(7):
        p = zu.copy()
        p[k] += 1j * hstep
        _, J = rk4(p[:5], p[5:], dt, wb)
        H[:, k] = np.einsum("rn,rjn->jn", y, J).imag / hstep
    return 0.5 * (H + H.transpose(1, 0, 2))


class MpcNLP:
    """prob: dict(N, dt, wb, obs_A (O,4,2), obs_b (O,4), body_G, body_g, region, limits, dmin, n_others)
    params: dict(cur (5,), ref (N,3), others (Vo,N,3))."""

    def __init__(self, prob, params, rho=1e3):
        self.p, self.par, self.rho = prob, params, float(rho)
        N, O, Vo = prob["N"], prob["obs_A"].shape[0], prob["n_others"]
        self.N, self.O, self.Vo = N, O, Vo
        va = _Alloc()
        self.iz = va.take(N, 7)
        self.ilam, self.imu = va.take(N, O, 4), va.take(N, O, 4)
        self.isd, self.iel = va.take(N, O), va.take(N, O)
        self.ipl, self.ipm, self.ips = va.take(Vo, N, 4), va.take(Vo, N, 4), va.take(Vo, N, 2)
        self.ipsd, self.ipsn, self.ipel = va.take(Vo, N), va.take(Vo, N), va.take(Vo, N)
        self.n = va.n
        xL, xU = np.full(self.n, -INF), np.full(self.n, INF)
        rg, lm = prob["region"], prob["limits"]
        xL[self.iz] = [rg[0], rg[2], -INF, lm[0], lm[2], lm[4], lm[6]]
        xU[self.iz] = [rg[1], rg[3], INF, lm[1], lm[3], lm[5], lm[7]]
        for arr in (self.ilam, self.imu, self.isd, self.iel, self.ipl, self.ipm, self.ipsd, self.ipsn, self.ipel):
            xL[arr] = 0.0
        self.xL, self.xU = xL, xU
        ra = _Alloc()
        self.r_init = ra.take(5)
        self.r_dyn = ra.take(N - 1, 5)
        self.r_obs = [ra.take(N, O, 4)]  # lists: same attribute shape as CollocationNLP (delta_c rows)
        self.r_pair = [ra.take(N, 6) for _ in range(Vo)]
        self.m = ra.n
        self.clip_pos = self.r_obs[0][:, :, 3].ravel()
        self.clip_neg = np.concatenate([r[:, 5] for r in self.r_pair]) if Vo else np.zeros(0, dtype=int)
        self.blks = []
        for j in range(O):
            loc = np.concatenate([self.iz[:, :3], self.ilam[:, j], self.imu[:, j], self.isd[:, j : j + 1], self.iel[:, j : j + 1]], axis=1)
            par = np.tile(np.concatenate([prob["obs_A"][j].ravel(), prob["obs_b"][j], prob["body_G"].ravel(), prob["body_g"], [prob["dmin"]]]), (N, 1))
            self.blks.append((blocks.obs_block(), loc, self.r_obs[0][:, j], par))
        for o in range(Vo):
            loc = np.concatenate([self.iz[:, :3], self.ipl[o], self.ipm[o], self.ips[o], self.ipsd[o][:, None], self.ipsn[o][:, None], self.ipel[o][:, None]], axis=1)
            par = np.concatenate([np.tile(np.concatenate([prob["body_G"].ravel(), prob["body_g"], [prob["dmin"]]]), (N, 1)), params["others"][o]], axis=1)
            self.blks.append((blocks.pair_block(other_is_param=True), loc, self.r_pair[o], par))

    def _elastic(self):
        return np.concatenate([self.iel.ravel(), self.ipel.ravel()])

    def f(self, x):
        z = x[self.iz]
        ref = self.par["ref"]
        cost = 100 * ((z[:, 0] - ref[:, 0]) ** 2 + (z[:, 1] - ref[:, 1]) ** 2 + (z[:, 2] - ref[:, 2]) ** 2)
        cost = cost + z[:, 5] ** 2 + z[:, 3] ** 2 * z[:, 6] ** 2 + z[:, 4] ** 2
        return float(cost.sum() + self.rho * x[self._elastic()].sum())

    def grad_f(self, x):
        g = np.zeros(self.n)
        z = x[self.iz]
        ref = self.par["ref"]
        g[self.iz[:, 0]] = 200 * (z[:, 0] - ref[:, 0])
        g[self.iz[:, 1]] = 200 * (z[:, 1] - ref[:, 1])
        g[self.iz[:, 2]] = 200 * (z[:, 2] - ref[:, 2])
        g[self.iz[:, 3]] = 2 * z[:, 3] * z[:, 6] ** 2
        g[self.iz[:, 4]] = 2 * z[:, 4]
        g[self.iz[:, 5]] = 2 * z[:, 5]
        g[self.iz[:, 6]] = 2 * z[:, 3] ** 2 * z[:, 6]
        g[self._elastic()] = self.rho
        return g

    def c(self, x):
        out = np.zeros(self.m)
        z = x[self.iz]
        out[self.r_init] = z[0, :5] - self.par["cur"]
        F, _ = rk4(z[:-1, :5].T.copy(), z[:-1, 5:].T.copy(), self.p["dt"], self.p["wb"])
        out[self.r_dyn] = z[1:, :5] - F.T
        for blk, loc, rows, par in self.blks:
            out[rows] = blk.c(x[loc].T, par.T).T
        return out

    def jac(self, x):
        rr, cc, vv = [], [], []
        z = x[self.iz]
        rr.append(self.r_init), cc.append(self.iz[0, :5]), vv.append(np.ones(5))
        _, dF = rk4(z[:-1, :5].T.copy(), z[:-1, 5:].T.copy(), self.p["dt"], self.p["wb"])  # (5,7,N-1)
        for r in range(5):
            rr.append(self.r_dyn[:, r]), cc.append(self.iz[1:, r]), vv.append(np.ones(self.N - 1))
            for k in range(7):
                rr.append(self.r_dyn[:, r]), cc.append(self.iz[:-1, k]), vv.append(-dF[r, k])
        for blk, loc, rows, par in self.blks:
            J = blk.jac(x[loc].T, par.T)
            for (r, c), v in zip(blk.jac_pat, J):
                rr.append(rows[:, r]), cc.append(loc[:, c]), vv.append(v)
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(self.m, self.n))

    def hess(self, x, y, clip=True):
        y = np.array(y, dtype=float)
        if clip:
            y[self.clip_pos] = np.maximum(y[self.clip_pos], 0.0)
            if len(self.clip_neg):
                y[self.clip_neg] = np.minimum(y[self.clip_neg], 0.0)
        rr, cc, vv = [], [], []
        z = x[self.iz]
        N = self.N
        for k, val in ((0, 200.0), (1, 200.0), (2, 200.0), (4, 2.0), (5, 2.0)):
            rr.append(self.iz[:, k]), cc.append(self.iz[:, k]), vv.append(np.full(N, val))
        rr.append(self.iz[:, 3]), cc.append(self.iz[:, 3]), vv.append(2 * z[:, 6] ** 2)
        rr.append(self.iz[:, 6]), cc.append(self.iz[:, 6]), vv.append(2 * z[:, 3] ** 2)
        for a, b in ((3, 6), (6, 3)):
            rr.append(self.iz[:, a]), cc.append(self.iz[:, b]), vv.append(4 * z[:, 3] * z[:, 6])
        H = rk4_hess(z[:-1, :5].T.copy(), z[:-1, 5:].T.copy(), y[self.r_dyn].T, self.p["dt"], self.p["wb"])  # (7,7,N-1)
        for a in range(7):
            for b in range(7):
                rr.append(self.iz[:-1, a]), cc.append(self.iz[:-1, b]), vv.append(-H[a, b])
        for blk, loc, rows, par in self.blks:
            Hb = blk.hes(x[loc].T, par.T, y[rows].T)
            for (r, c), v in zip(blk.hes_pat, Hb):
                rr.append(loc[:, r]), cc.append(loc[:, c]), vv.append(v)
                if r != c:
                    rr.append(loc[:, c]), cc.append(loc[:, r]), vv.append(v)
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(self.n, self.n))

    def init_slacks(self, x):
        x = x.copy()
        sl = np.concatenate([self.isd.ravel(), self.ipsd.ravel(), self.ipsn.ravel()])
        el = self._elastic()
        x[sl] = 0.0
        x[el] = 0.0
        body = self.c(x)
        d_rows = np.concatenate([self.r_obs[0][:, :, 0].ravel()] + [r[:, 0] for r in self.r_pair])
        d_slk = np.concatenate([self.isd.ravel(), self.ipsd.ravel()])
        x[d_slk] = np.maximum(body[d_rows], 0.0)
        x[el] = np.maximum(-body[d_rows], 0.0)
        if self.Vo:
            x[self.ipsn.ravel()] = body[np.concatenate([r[:, 5] for r in self.r_pair])]
        return x

    def pack(self, z, lam, mu, pl=None, pm=None, ps=None):
        x = np.zeros(self.n)
        x[self.iz], x[self.ilam], x[self.imu] = z, lam, mu
        if self.Vo and pl is not None:
            x[self.ipl], x[self.ipm], x[self.ips] = pl, pm, ps
        return x
