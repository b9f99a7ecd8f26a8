"""ORACLE (test infrastructure, never imported by the product path).

Sparse restatement of the collocation OBCA NLP: single vehicle
(confrez/control/vehicle.py:360-640) and joint (multi_vehicle_planner.py:343-480), assembled
from the sympy blocks in :mod:`oracle.blocks`.

    min f(x)   s.t.  c(x) = 0,  xL <= x <= xU

with every reference inequality rewritten as ``g(x) - s = 0, s >= 0`` (IPOPT's own slack form).
Variable / row numbering is private to the oracle; ``unpack``/``pack`` convert to the named
arrays of ``conflict_rez_b200.problem`` (z, lam, mu, pair duals, dt) that the C ABI exchanges.

Parity status: the reference ships no golden vectors for this path and CasADi/IPOPT are not
installable here => **parity unpinned** at the CasADi boundary.  The pins this repo creates are
finite-difference checks of every block and KKT-residual checks of every solution
(tests/test_oracle_*.py).
"""
from itertools import combinations

import numpy as np
import scipy.sparse as sp

from oracle import blocks
from oracle.collocation import collocation_coefficients

INF = np.inf


class _Alloc:
    def __init__(self):
        self.n = 0

    def take(self, *shape):
        cnt = int(np.prod(shape)) if shape else 1
        idx = np.arange(self.n, self.n + cnt).reshape(shape)
        self.n += cnt
        return idx


class CollocationNLP:
    def __init__(self, prob, rho=1e3):
        p = self.prob = prob
        self.rho = float(rho)  # weight of the exact l1 penalty on the elastic variables
        assert p.batch is None, "the oracle works on one instance; use prob.instance(b)"
        K = self.K = p.K
        V, O = p.V, p.O
        self.N = [int(n) for n in p.N]
        self.M = [int(m) for m in p.nodes]
        self.A, self.B, self.D = collocation_coefficients(K)
        self.pairs = list(combinations(range(V), 2))
        self.Mp = [min(self.M[a], self.M[b]) for a, b in self.pairs]

        # ---------------- variables
        va = _Alloc()
        self.iz = [va.take(self.M[a], 7) for a in range(V)]
        self.ilam = [va.take(self.M[a], O, 4) for a in range(V)]
        self.imu = [va.take(self.M[a], O, 4) for a in range(V)]
        self.isd = [va.take(self.M[a], O) for a in range(V)]
        self.iel = [va.take(self.M[a], O) for a in range(V)]
        self.its = [va.take(int(p.n_sets[a]) - 1, 8) for a in range(V)]
        self.ipl = [va.take(m, 4) for m in self.Mp]
        self.ipm = [va.take(m, 4) for m in self.Mp]
        self.ips = [va.take(m, 2) for m in self.Mp]
        self.ipsd = [va.take(m) for m in self.Mp]
        self.ipsn = [va.take(m) for m in self.Mp]
        self.ipel = [va.take(m) for m in self.Mp]
        self.idt = int(va.take())
        self.n = va.n

        xL = np.full(self.n, -INF)
        xU = np.full(self.n, INF)
        lo = [p.region[0], p.region[2], -INF, p.limits[0], p.limits[2], p.limits[4], p.limits[6]]
        hi = [p.region[1], p.region[3], INF, p.limits[1], p.limits[3], p.limits[5], p.limits[7]]
        for a in range(V):
            xL[self.iz[a]] = lo
            xU[self.iz[a]] = hi
            for arr in (self.ilam[a], self.imu[a], self.isd[a], self.iel[a], self.its[a]):
                xL[arr] = 0.0
        for q in range(len(self.pairs)):
            for arr in (self.ipl[q], self.ipm[q], self.ipsd[q], self.ipsn[q], self.ipel[q]):
                xL[arr] = 0.0
        self.xL, self.xU = xL, xU

        # ---------------- constraint rows
        ra = _Alloc()
        self.r_init = [ra.take(7) for a in range(V)]
        self.r_col = [ra.take(self.M[a], 5) for a in range(V)]
        self.r_cont = [ra.take(self.N[a] - 1, 7) for a in range(V)]
        self.has_heading = [bool(np.isfinite(p.final_heading[a])) for a in range(V)]
        self.r_term = [ra.take(5 if self.has_heading[a] else 4) for a in range(V)]
        self.r_obs = [ra.take(self.M[a], O, 4) for a in range(V)]
        self.r_tube = [ra.take(int(p.n_sets[a]) - 1, 8) for a in range(V)]
        self.r_pair = [ra.take(m, 6) for m in self.Mp]
        self.m = ra.n

        self._build_linear()
        self._build_blocks()

    # ------------------------------------------------------------------ linear rows
    def _build_linear(self):
        p, K, D = self.prob, self.K, self.D
        rows, cols, vals = [], [], []
        rhs = np.zeros(self.m)

        def add(r, c, v):
            rows.append(np.ravel(r)), cols.append(np.ravel(c)), vals.append(np.broadcast_to(v, np.shape(np.ravel(r))).astype(float))

        for a in range(p.V):
            iz, N = self.iz[a], self.N[a]
            # initial state (vehicle.py:424-434)
            add(self.r_init[a], iz[0], 1.0)
            rhs[self.r_init[a]] = [p.init_pose[a, 0], p.init_pose[a, 1], p.init_pose[a, 2], 0, 0, 0, 0]
            # continuity (vehicle.py:544-568): sum_j D_j z_{i-1,j} - z_{i,0} = 0  (states and inputs)
            for i in range(1, N):
                r = self.r_cont[a][i - 1]
                for j in range(K + 1):
                    add(r, iz[(K + 1) * (i - 1) + j], D[j])
                add(r, iz[(K + 1) * i], -1.0)
            # terminal (vehicle.py:619-626): zF_psi = heading, zF_v = zF_delta = 0, uF = 0
            comps = ([2] if self.has_heading[a] else []) + [3, 4, 5, 6]
            for r, comp in zip(self.r_term[a], comps):
                for j in range(K + 1):
                    add(r, iz[(K + 1) * (N - 1) + j, comp], D[j])
            if self.has_heading[a]:
                rhs[self.r_term[a][0]] = p.final_heading[a]
        self.A_lin = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(self.m, self.n))
        self.rhs_lin = rhs

    # ------------------------------------------------------------------ nonlinear blocks
    def _build_blocks(self):
        p, K = self.prob, self.K
        V, O = p.V, p.O
        self.blks = []  # (block, loc_idx (n,nloc), row_idx (n,nrow) or None for the objective, params (n,npar))

        def tube_par(a, q):
            return np.concatenate(
                [p.tube_A[a, q, 0].ravel(), p.tube_b[a, q, 0] - p.shrink_tube, p.tube_A[a, q, 1].ravel(), p.tube_b[a, q, 1] - p.shrink_tube, [p.wb]]
            )

        for a in range(V):
            iz, N, M = self.iz[a], self.N[a], self.M[a]
            node0 = (K + 1) * np.arange(N)
            # collocation (vehicle.py:487-509) at every k = 0..K
            for k in range(K + 1):
                zall = np.stack([iz[node0 + j, :5] for j in range(K + 1)], axis=1).reshape(N, -1)
                loc = np.concatenate([zall, iz[node0 + k, 5:7], np.full((N, 1), self.idt)], axis=1)
                par = np.tile(np.append(self.A[:, k], p.wb), (N, 1))
                self.blks.append((blocks.col_block(k, K), loc, self.r_col[a][node0 + k], par))
            # running cost (vehicle.py:512-521)
            loc = np.concatenate([iz[:, 3:7], np.full((M, 1), self.idt)], axis=1)
            par = np.tile(self.B, N)[:, None]
            self.blks.append((blocks.cost_block(), loc, None, par))
            # obstacles (vehicle.py:524-541)
            for j in range(O):
                loc = np.concatenate([iz[:, :3], self.ilam[a][:, j], self.imu[a][:, j], self.isd[a][:, j : j + 1], self.iel[a][:, j : j + 1]], axis=1)
                par = np.tile(np.concatenate([p.obs_A[j].ravel(), p.obs_b[j], p.body_G.ravel(), p.body_g, [p.dmin]]), (M, 1))
                self.blks.append((blocks.obs_block(), loc, self.r_obs[a][:, j], par))
            # tube sets at set transitions (vehicle.py:570-584): q = 1..S-2 at node (q*n_per_set, 0)
            S = int(p.n_sets[a])
            if S > 2:
                qs = np.arange(1, S - 1)
                loc = np.concatenate([iz[(K + 1) * p.n_per_set * qs, :3], self.its[a][qs - 1]], axis=1)
                par = np.stack([tube_par(a, q) for q in qs])
                self.blks.append((blocks.tube_block(), loc, self.r_tube[a][qs - 1], par))
            # tube on the end state (vehicle.py:605-617) with the last set
            last = (K + 1) * (N - 1) + np.arange(K + 1)
            loc = np.concatenate([iz[last, :3].ravel(), self.its[a][S - 2]])[None, :]
            par = np.concatenate([tube_par(a, S - 1), self.D])[None, :]
            self.blks.append((blocks.tubeF_block(K), loc, self.r_tube[a][S - 2][None, :], par))
        # vehicle pairs (multi_vehicle_planner.py:419-451)
        for q, (a, b) in enumerate(self.pairs):
            m = self.Mp[q]
            loc = np.concatenate(
                [self.iz[a][:m, :3], self.iz[b][:m, :3], self.ipl[q], self.ipm[q], self.ips[q], self.ipsd[q][:, None], self.ipsn[q][:, None], self.ipel[q][:, None]], axis=1
            )
            par = np.tile(np.concatenate([p.body_G.ravel(), p.body_g, [p.dmin]]), (m, 1))
            self.blks.append((blocks.pair_block(), loc, self.r_pair[q], par))

        # rows whose multiplier is clipped in the Hessian (algorithm spec, DESIGN.md "local convexification"):
        # obstacle norm row c3 uses max(y,0); pair norm row uses min(y,0) (both exact at any KKT point).
        self.clip_pos = np.concatenate([self.r_obs[a][:, :, 3].ravel() for a in range(V)])
        self.clip_neg = np.concatenate([r[:, 5].ravel() for r in self.r_pair]) if self.pairs else np.zeros(0, dtype=int)

    # ------------------------------------------------------------------ evaluation
    def _elastic_index(self):
        return np.concatenate([a.ravel() for a in self.iel] + [a.ravel() for a in self.ipel])

    def f(self, x):
        tot = sum((self.N[a] * x[self.idt]) ** 2 for a in range(self.prob.V))
        tot += self.rho * x[self._elastic_index()].sum()
        for blk, loc, rows, par in self.blks:
            if rows is None:
                tot += blk.c(x[loc].T, par.T).sum()
        return float(tot)

    def grad_f(self, x):
        g = np.zeros(self.n)
        g[self._elastic_index()] = self.rho
        g[self.idt] += sum(2 * self.N[a] ** 2 * x[self.idt] for a in range(self.prob.V))
        for blk, loc, rows, par in self.blks:
            if rows is None:
                J = blk.jac(x[loc].T, par.T)
                for (r, c), v in zip(blk.jac_pat, J):
                    np.add.at(g, loc[:, c], v)
        return g

    def c(self, x):
        out = self.A_lin @ x - self.rhs_lin
        for blk, loc, rows, par in self.blks:
            if rows is not None:
                out[rows] = blk.c(x[loc].T, par.T).T
        return out

    def jac(self, x):
        rr, cc, vv = [], [], []
        for blk, loc, rows, par in self.blks:
            if rows is None:
                continue
            J = blk.jac(x[loc].T, par.T)
            for (r, c), v in zip(blk.jac_pat, J):
                rr.append(rows[:, r]), cc.append(loc[:, c]), vv.append(v)
        Jn = sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(self.m, self.n))
        return (Jn + self.A_lin).tocsr()

    def hess(self, x, y, clip=True):
        """Full symmetric Hessian of f + y'c.  ``clip`` applies the local convexification of the algorithm spec."""
        y = np.array(y, dtype=float)
        if clip:
            y[self.clip_pos] = np.maximum(y[self.clip_pos], 0.0)
            if len(self.clip_neg):
                y[self.clip_neg] = np.minimum(y[self.clip_neg], 0.0)
        rr, cc, vv = [], [], []
        for blk, loc, rows, par in self.blks:
            n = loc.shape[0]
            Y = np.ones((1, n)) if rows is None else y[rows].T
            H = blk.hes(x[loc].T, par.T, Y)
            for (r, c), v in zip(blk.hes_pat, H):
                rr.append(loc[:, r]), cc.append(loc[:, c]), vv.append(v)
                if r != c:
                    rr.append(loc[:, c]), cc.append(loc[:, r]), vv.append(v)
        rr.append(np.array([self.idt])), cc.append(np.array([self.idt]))
        vv.append(np.array([sum(2.0 * self.N[a] ** 2 for a in range(self.prob.V))]))
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(self.n, self.n))

    # ------------------------------------------------------------------ named views
    def pack(self, guess):
        """CollocationGuess -> flat x (slacks are left at 0; the IPM initialises them from the constraints)."""
        x = np.zeros(self.n)
        for a in range(self.prob.V):
            M = self.M[a]
            x[self.iz[a]] = guess.z[a, :M]
            x[self.ilam[a]] = guess.lam[a, :M]
            x[self.imu[a]] = guess.mu[a, :M]
        for q in range(len(self.pairs)):
            m = self.Mp[q]
            if guess.pair_lam is not None:
                x[self.ipl[q]] = guess.pair_lam[q, :m]
                x[self.ipm[q]] = guess.pair_mu[q, :m]
                x[self.ips[q]] = guess.pair_s[q, :m]
        x[self.idt] = float(guess.dt)
        return x

    def slack_index(self):
        """Indices of all slack variables and of the rows that define them (slack = row residual at slack 0)."""
        idx, rows = [], []
        for a in range(self.prob.V):
            idx += [self.isd[a].ravel(), self.its[a].ravel()]
            rows += [self.r_obs[a][:, :, 0].ravel(), self.r_tube[a].ravel()]
        for q in range(len(self.pairs)):
            idx += [self.ipsd[q], self.ipsn[q]]
            rows += [self.r_pair[q][:, 0], self.r_pair[q][:, 5]]
        return np.concatenate(idx), np.concatenate(rows)

    def init_slacks(self, x):
        """Set every slack to the value of its inequality body (IPOPT: s0 = g(x0)); on elastic rows the
        negative part of the body goes to the elastic variable so that the row starts feasible."""
        x = x.copy()
        idx, rows = self.slack_index()
        el = self._elastic_index()
        x[idx] = 0.0
        x[el] = 0.0
        body = self.c(x)
        x[idx] = body[rows]
        erow = np.concatenate([self.r_obs[a][:, :, 0].ravel() for a in range(self.prob.V)] + [r[:, 0] for r in self.r_pair])
        eslk = np.concatenate([self.isd[a].ravel() for a in range(self.prob.V)] + [i for i in self.ipsd])
        x[eslk] = np.maximum(body[erow], 0.0)
        x[el] = np.maximum(-body[erow], 0.0)
        return x

    def unpack(self, x):
        p = self.prob
        Mmax = max(self.M)
        out = {
            "z": np.zeros((p.V, Mmax, 7)),
            "lam": np.zeros((p.V, Mmax, p.O, 4)),
            "mu": np.zeros((p.V, Mmax, p.O, 4)),
            "dt": float(x[self.idt]),
        }
        for a in range(p.V):
            M = self.M[a]
            out["z"][a, :M] = x[self.iz[a]]
            out["lam"][a, :M] = x[self.ilam[a]]
            out["mu"][a, :M] = x[self.imu[a]]
        P = len(self.pairs)
        out["pair_lam"] = np.zeros((P, Mmax, 4))
        out["pair_mu"] = np.zeros((P, Mmax, 4))
        out["pair_s"] = np.zeros((P, Mmax, 2))
        for q in range(P):
            m = self.Mp[q]
            out["pair_lam"][q, :m] = x[self.ipl[q]]
            out["pair_mu"][q, :m] = x[self.ipm[q]]
            out["pair_s"][q, :m] = x[self.ips[q]]
        return out
