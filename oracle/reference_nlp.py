"""ORACLE (test infrastructure, never imported by the product path).

The UNMODIFIED reference NLP and an algorithm-independent KKT certificate for it.

``oracle/nlp.py`` + ``oracle/ipm.py`` restate the problem in the form the solvers iterate on (slack variables, elastic
variables on the distance rows, clipped norm-row multipliers in the Hessian) -- modifications the CUDA solver shares.
This module has none of them.  It writes the reference's constraint list down literally, with hard inequalities:

    confrez/control/vehicle.py:424-434   initial state          (equalities)
    confrez/control/vehicle.py:437-485   simple bounds, l >= 0, m >= 0
    confrez/control/vehicle.py:487-509   collocation, every k = 0..K        (equalities)
    confrez/control/vehicle.py:511-521   running cost, 638 the (N dt)^2 term
    confrez/control/vehicle.py:523-541   OBCA: dist >= dmin, G'm + R'A'l = 0, |A'l|^2 = 1
    confrez/control/vehicle.py:544-568   continuity of states and inputs    (equalities)
    confrez/control/vehicle.py:570-584   tube sets at the set transitions   (inequalities)
    confrez/control/vehicle.py:586-626   end state: tube set, heading, v = delta = 0, uF = 0
    confrez/control/multi_vehicle_planner.py:419-451   vehicle pairs: -b_i'l - b_j'm >= dmin, A_i'l + s = 0, A_j'm - s = 0, s's <= 1

as differentiable torch (float64, CPU) functions; first derivatives come from torch.autograd -- a third derivative path
next to sympy (oracle/blocks.py) and the hand-written CUDA.  No second derivatives, no barrier, no line search: the
certificate only asks whether a returned point, with some multipliers, is a first-order KKT point of the reference problem:

    stationarity   grad f + J' y - zL + zU = 0
    feasibility    h = 0, g >= 0, lo <= w <= hi
    signs          y <= 0 on inequality rows (L = f + y'c, rows written as c >= 0), zL, zU >= 0
    complementarity  y_i g_i = 0, zL (w - lo) = 0, zU (hi - w) = 0

Multipliers are supplied (a solver's own; whoever produced them, multipliers that satisfy the conditions prove the KKT
property) or, on small problems, recovered here from nothing by a bounded least-squares fit (``recover=True``).
"""
from itertools import combinations

import numpy as np
import torch

from oracle.collocation import collocation_coefficients

T = torch.float64
EQ_ROWS = ("init", "col", "cont", "term", "obs_rot", "obs_norm", "pair_e1", "pair_e2")
INEQ_ROWS = ("obs_dist", "tube", "pair_dist", "pair_norm")


def _t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=T)


class ReferenceNLP:
    """Reference collocation OBCA NLP of one instance (``prob.batch is None``)."""

    def __init__(self, prob):
        assert prob.batch is None
        self.p = prob
        self.V, self.O, self.K = prob.V, prob.O, prob.K
        self.N = [int(n) for n in prob.N]
        self.M = [int(m) for m in prob.nodes]
        A, B, D = collocation_coefficients(prob.K)
        self.A, self.B, self.D = _t(A), _t(B), _t(D)
        self.pairs = list(combinations(range(self.V), 2))
        self.Mp = [min(self.M[a], self.M[b]) for a, b in self.pairs]
        self.G, self.g = _t(prob.body_G), _t(prob.body_g)
        self.obsA, self.obsb = _t(prob.obs_A), _t(prob.obs_b)
        lo = [prob.region[0], prob.region[2], -np.inf, prob.limits[0], prob.limits[2], prob.limits[4], prob.limits[6]]
        hi = [prob.region[1], prob.region[3], np.inf, prob.limits[1], prob.limits[3], prob.limits[5], prob.limits[7]]
        self.zlo, self.zhi = np.array(lo), np.array(hi)

    # ---------------------------------------------------------------- variables
    def variables(self, z, lam, mu, dt, pair_lam=None, pair_mu=None, pair_s=None, requires_grad=True):
        """Named reference variables from the arrays the C ABI exchanges (padding nodes dropped)."""
        w = {}
        for a in range(self.V):
            M = self.M[a]
            w["z%d" % a] = _t(np.asarray(z)[a, :M])
            w["l%d" % a] = _t(np.asarray(lam)[a, :M])
            w["m%d" % a] = _t(np.asarray(mu)[a, :M])
        for q, m in enumerate(self.Mp):
            w["pl%d" % q] = _t(np.asarray(pair_lam)[q, :m])
            w["pm%d" % q] = _t(np.asarray(pair_mu)[q, :m])
            w["ps%d" % q] = _t(np.asarray(pair_s)[q, :m])
        w["dt"] = _t(float(np.asarray(dt)))
        if requires_grad:
            for v in w.values():
                v.requires_grad_(True)
        return w

    def bounds(self, w):
        lo, hi = {}, {}
        for k, v in w.items():
            lo[k] = torch.full_like(v, -np.inf)
            hi[k] = torch.full_like(v, np.inf)
            if k[0] == "z":
                lo[k] = _t(np.broadcast_to(self.zlo, tuple(v.shape)).copy())
                hi[k] = _t(np.broadcast_to(self.zhi, tuple(v.shape)).copy())
            elif k[0] in "lm" or k[:2] in ("pl", "pm"):
                lo[k] = torch.zeros_like(v)
        return lo, hi

    # ---------------------------------------------------------------- objective and constraints
    def objective(self, w):
        dt = w["dt"]
        J = 0.0
        for a in range(self.V):
            z = w["z%d" % a].reshape(self.N[a], self.K + 1, 7)
            err = z[..., 5] ** 2 + z[..., 3] ** 2 * z[..., 6] ** 2 + z[..., 4] ** 2      # vehicle.py:512-520
            J = J + (self.B[None, :] * err).sum() * dt + (self.N[a] * dt) ** 2            # vehicle.py:521, 638
        return J

    def constraints(self, w):
        """dict name -> tensor; equality rows (EQ_ROWS) are = 0, inequality rows (INEQ_ROWS) are >= 0."""
        p, K, dt = self.p, self.K, w["dt"]
        out = {}
        for a in range(self.V):
            N, M = self.N[a], self.M[a]
            z = w["z%d" % a]
            zi = z.reshape(N, K + 1, 7)
            init = torch.cat([_t(p.init_pose[a]), torch.zeros(4, dtype=T)])
            out["init%d" % a] = z[0] - init                                                # vehicle.py:424-434
            x, y, psi, v, de, ua, uw = [zi[..., c] for c in range(7)]
            f = torch.stack([v * torch.cos(psi), v * torch.sin(psi), v / p.wb * torch.tan(de), ua, uw], -1)
            poly = torch.einsum("jk,ijc->ikc", self.A, zi[..., :5]) / dt                   # vehicle.py:499-507
            out["col%d" % a] = (poly - f).reshape(M, 5)                                    # vehicle.py:509
            zend = torch.einsum("j,ijc->ic", self.D, zi)                                   # sum_j D_j z_{i,j}  (states and inputs)
            out["cont%d" % a] = zend[:-1] - zi[1:, 0]                                       # vehicle.py:544-568
            zF = zend[-1]
            term = [zF[3], zF[4], zF[5], zF[6]]                                             # vehicle.py:622-626
            if np.isfinite(p.final_heading[a]):
                term = [zF[2] - float(p.final_heading[a])] + term                          # vehicle.py:619-620
            out["term%d" % a] = torch.stack(term)
            # OBCA rows, every node and obstacle (vehicle.py:523-541)
            l, m = w["l%d" % a], w["m%d" % a]                                               # (M, O, 4)
            t = z[:, :2]
            c, s = torch.cos(z[:, 2]), torch.sin(z[:, 2])
            Atb = torch.einsum("orx,nx->nor", self.obsA, t) - self.obsb[None]               # obs.A @ t - obs.b
            u = torch.einsum("orx,nor->nox", self.obsA, l)                                  # obs.A.T @ lj
            out["obs_dist%d" % a] = -(m * self.g).sum(-1) + (Atb * l).sum(-1) - p.dmin       # >= 0
            Gm = torch.einsum("rx,nor->nox", self.G, m)                                     # veh_G.T @ mj
            Rtu = torch.stack([c[:, None] * u[..., 0] + s[:, None] * u[..., 1], -s[:, None] * u[..., 0] + c[:, None] * u[..., 1]], -1)
            out["obs_rot%d" % a] = Gm + Rtu                                                # = 0
            out["obs_norm%d" % a] = (u * u).sum(-1) - 1.0                                   # = 0
            # tube sets (vehicle.py:570-584 at i = q N_per_set, k = 0; 605-617 at the end state with the last set)
            S = int(p.n_sets[a])
            rows = []
            for q in range(1, S):
                pose = zF if q == S - 1 else zi[q * p.n_per_set, 0]
                back = pose[:2]
                front = torch.stack([pose[0] + p.wb * torch.cos(pose[2]), pose[1] + p.wb * torch.sin(pose[2])])
                rows.append(_t(p.tube_b[a, q, 0]) - p.shrink_tube - _t(p.tube_A[a, q, 0]) @ back)
                rows.append(_t(p.tube_b[a, q, 1]) - p.shrink_tube - _t(p.tube_A[a, q, 1]) @ front)
            out["tube%d" % a] = torch.cat(rows).reshape(S - 1, 8)                           # >= 0
        for q, (a, b) in enumerate(self.pairs):                                             # multi_vehicle_planner.py:419-451
            m = self.Mp[q]
            za, zb = w["z%d" % a][:m], w["z%d" % b][:m]
            lik, mik, sik = w["pl%d" % q], w["pm%d" % q], w["ps%d" % q]

            def body(zz):
                c, s = torch.cos(-zz[:, 2]), torch.sin(-zz[:, 2])
                R = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)      # R(-psi)
                Av = torch.einsum("rx,nxy->nry", self.G, R)                                   # veh_G @ R
                bv = torch.einsum("nry,ny->nr", Av, zz[:, :2]) + self.g                       # veh_G @ R @ t + veh_g
                return Av, bv

            Aa, ba = body(za)
            Ab, bb = body(zb)
            out["pair_dist%d" % q] = -(ba * lik).sum(-1) - (bb * mik).sum(-1) - p.dmin       # >= 0
            out["pair_e1%d" % q] = torch.einsum("nry,nr->ny", Aa, lik) + sik                  # = 0
            out["pair_e2%d" % q] = torch.einsum("nry,nr->ny", Ab, mik) - sik                  # = 0
            out["pair_norm%d" % q] = 1.0 - (sik * sik).sum(-1)                               # >= 0
        return out

    @staticmethod
    def is_inequality(name):
        return name.rstrip("0123456789") in INEQ_ROWS

    # ---------------------------------------------------------------- multipliers of an oracle / solver iterate
    def multipliers_from_oracle(self, nlp, y, zL, zU):
        """(y, zL, zU) in the ordering of ``oracle.nlp.CollocationNLP`` -> named multipliers of the reference problem.
        Inequality rows of the slack form ``g - s (+ e) = 0`` carry the same multiplier in ``L = f + y'c``."""
        y, zL, zU = np.asarray(y), np.asarray(zL), np.asarray(zU)
        my, mL, mU = {}, {}, {}
        for a in range(self.V):
            my["init%d" % a] = _t(y[nlp.r_init[a]])
            my["col%d" % a] = _t(y[nlp.r_col[a]])
            my["cont%d" % a] = _t(y[nlp.r_cont[a]])
            my["term%d" % a] = _t(y[nlp.r_term[a]])
            ro = nlp.r_obs[a]
            my["obs_dist%d" % a] = _t(y[ro[:, :, 0]])
            my["obs_rot%d" % a] = _t(y[ro[:, :, 1:3]])
            my["obs_norm%d" % a] = _t(y[ro[:, :, 3]])
            my["tube%d" % a] = _t(y[nlp.r_tube[a]])
            for nm, idx in (("z", nlp.iz[a]), ("l", nlp.ilam[a]), ("m", nlp.imu[a])):
                mL["%s%d" % (nm, a)], mU["%s%d" % (nm, a)] = _t(zL[idx]), _t(zU[idx])
        for q in range(len(self.pairs)):
            rp = nlp.r_pair[q]
            my["pair_dist%d" % q] = _t(y[rp[:, 0]])
            my["pair_e1%d" % q] = _t(y[rp[:, 1:3]])
            my["pair_e2%d" % q] = _t(y[rp[:, 3:5]])
            my["pair_norm%d" % q] = _t(y[rp[:, 5]])
            for nm, idx in (("pl", nlp.ipl[q]), ("pm", nlp.ipm[q]), ("ps", nlp.ips[q])):
                mL["%s%d" % (nm, q)], mU["%s%d" % (nm, q)] = _t(zL[idx]), _t(zU[idx])
        mL["dt"], mU["dt"] = _t(zL[nlp.idt]), _t(zU[nlp.idt])
        return my, mL, mU


def _flat(d, keys):
    return torch.cat([d[k].reshape(-1) for k in keys])


def kkt_certificate(ref: ReferenceNLP, w, mult=None, recover=False):
    """First-order KKT residuals of the reference NLP at the point ``w`` (dict from ``ref.variables``).

    mult = (my, mL, mU): named multipliers -- a solver's own, mapped with ``multipliers_from_oracle``; any multipliers that
    satisfy the conditions prove that ``w`` is a KKT point, whoever produced them.  recover=True (small problems, dense
    algebra): the multipliers are computed here, from nothing, as the bounded least-squares solution of

        min | grad f + J' y - zL + zU |^2 + | g o y_ineq |^2 + | (w - lo) o zL |^2 + | (hi - w) o zU |^2
        s.t. y_ineq <= 0, zL >= 0, zU >= 0

    (stationarity and complementarity in one residual).  Returns a dict of max-norm residuals."""
    wk = list(w.keys())
    cons = ref.constraints(w)
    ck = list(cons.keys())
    f = ref.objective(w)
    gf = torch.autograd.grad(f, [w[k] for k in wk], retain_graph=True, allow_unused=True)
    gf = {k: (torch.zeros_like(w[k]) if g is None else g) for k, g in zip(wk, gf)}
    lo, hi = ref.bounds(w)

    def jt(my):
        """J' y over all variables (one reverse sweep)."""
        s = sum((cons[k] * my[k]).sum() for k in ck)
        g = torch.autograd.grad(s, [w[k] for k in wk], retain_graph=True, allow_unused=True)
        return {k: (torch.zeros_like(w[k]) if gi is None else gi) for k, gi in zip(wk, g)}

    if recover:
        from scipy.optimize import lsq_linear

        nvar = sum(int(w[k].numel()) for k in wk)
        assert nvar <= 4000, "recover=True uses dense algebra: small problems only"
        prim = tuple(w[k].detach() for k in wk)

        def fun(*args):
            c = ref.constraints(dict(zip(wk, args)))
            return torch.cat([c[k].reshape(-1) for k in ck])

        Jb = torch.autograd.functional.jacobian(fun, prim)
        J = torch.cat([j.reshape(j.shape[0], -1) for j in Jb], 1).numpy()          # (rows, nvar)
        cv = torch.cat([cons[k].reshape(-1) for k in ck]).detach().numpy()
        is_in = np.concatenate([np.full(int(cons[k].numel()), ref.is_inequality(k)) for k in ck])
        wv = torch.cat([w[k].reshape(-1) for k in wk]).detach().numpy()
        lov = torch.cat([lo[k].reshape(-1) for k in wk]).numpy()
        hiv = torch.cat([hi[k].reshape(-1) for k in wk]).numpy()
        iL, iU = np.flatnonzero(np.isfinite(lov)), np.flatnonzero(np.isfinite(hiv))
        m_rows = J.shape[0]
        A = np.zeros((nvar + int(is_in.sum()) + len(iL) + len(iU), m_rows + len(iL) + len(iU)))
        A[:nvar, :m_rows] = J.T
        A[iL, m_rows + np.arange(len(iL))] = -1.0
        A[iU, m_rows + len(iL) + np.arange(len(iU))] = 1.0
        r = nvar
        for j in np.flatnonzero(is_in):
            A[r, j] = cv[j]
            r += 1
        for q, j in enumerate(iL):
            A[r, m_rows + q] = wv[j] - lov[j]
            r += 1
        for q, j in enumerate(iU):
            A[r, m_rows + len(iL) + q] = hiv[j] - wv[j]
            r += 1
        rhs = np.concatenate([-torch.cat([gf[k].reshape(-1) for k in wk]).detach().numpy(), np.zeros(A.shape[0] - nvar)])
        lb = np.concatenate([np.full(m_rows, -np.inf), np.zeros(len(iL) + len(iU))])
        ub = np.concatenate([np.where(is_in, 0.0, np.inf), np.full(len(iL) + len(iU), np.inf)])
        sol = lsq_linear(A, rhs, bounds=(lb, ub), method="bvls", tol=1e-14, max_iter=20 * A.shape[1]).x
        my, o = {}, 0
        for k in ck:
            n = int(cons[k].numel())
            my[k], o = _t(sol[o:o + n]).reshape(cons[k].shape), o + n
        zLv, zUv = np.zeros(nvar), np.zeros(nvar)
        zLv[iL], zUv[iU] = sol[m_rows:m_rows + len(iL)], sol[m_rows + len(iL):]
        mL, mU, o = {}, {}, 0
        for k in wk:
            n = int(w[k].numel())
            mL[k], mU[k], o = _t(zLv[o:o + n]).reshape(w[k].shape), _t(zUv[o:o + n]).reshape(w[k].shape), o + n
    elif mult is None:
        raise ValueError("kkt_certificate needs multipliers (mult=...) or recover=True")
    else:
        my, mL, mU = mult

    g = jt(my)
    with torch.no_grad():
        stat = max(float((gf[k] + g[k] - mL[k] + mU[k]).abs().max()) for k in wk)
        eq = max(float(cons[k].abs().max()) for k in ck if not ref.is_inequality(k))
        ineq = max([float((-cons[k]).clamp(min=0).max()) for k in ck if ref.is_inequality(k)] + [0.0])
        bnd = max(float(torch.maximum(lo[k] - w[k], w[k] - hi[k]).clamp(min=0).max()) for k in wk)
        sign = max([float(my[k].clamp(min=0).max()) for k in ck if ref.is_inequality(k)] + [float((-mL[k]).clamp(min=0).max()) for k in wk] +
                   [float((-mU[k]).clamp(min=0).max()) for k in wk])
        compl = max([float((my[k] * cons[k]).abs().max()) for k in ck if ref.is_inequality(k)] +
                    [float(torch.where(torch.isfinite(lo[k]), mL[k] * (w[k] - lo[k]), torch.zeros_like(w[k])).abs().max()) for k in wk] +
                    [float(torch.where(torch.isfinite(hi[k]), mU[k] * (hi[k] - w[k]), torch.zeros_like(w[k])).abs().max()) for k in wk])
        tot = sum(float(my[k].abs().sum()) for k in ck) + sum(float(mL[k].sum() + mU[k].sum()) for k in wk)
        cnt = sum(int(my[k].numel()) for k in ck) + sum(int(torch.isfinite(lo[k]).sum() + torch.isfinite(hi[k]).sum()) for k in wk)
    s_d = max(100.0, tot / max(1, cnt)) / 100.0   # IPOPT's dual scaling
    return {"stationarity": stat, "stationarity_scaled": stat / s_d, "equality": eq, "inequality": ineq, "bounds": bnd, "sign": sign,
            "complementarity": compl, "objective": float(f.detach())}
