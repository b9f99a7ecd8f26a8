"""ORACLE (test infrastructure, never imported by the product path).

Plain sparse primal-dual interior-point method on the CPU, following the published IPOPT algorithm
(Waechter & Biegler 2006; option values as recalled in SURVEY.md App. E): monotone barrier, fraction to the
boundary, filter line search, inertia-correcting regularisation.  The reference reaches IPOPT through
``opti.solver("ipopt", ...)`` (confrez/control/vehicle.py:657, multi_vehicle_planner.py:464,
vehicle_follower.py:368); IPOPT itself is a third-party binary that is not in /root/reference and not
installable here, hence this restatement ("parity unpinned" at the IPOPT boundary).

The CUDA solver implements the same algorithm specification (DESIGN.md "IPM specification"), so iterates can be
compared step by step; the linear algebra differs completely (sparse LU here, structured elimination +
null-space Riccati there).

Differences from stock IPOPT (both sides of the parity test share them):
  * y0 = 0 (no least-squares multiplier estimate), no second-order correction, no restoration phase
    (a failed line search returns status ``Restoration_Failed``);
  * the Hessian uses clipped multipliers on the two norm rows (nlp.hess(clip=True));
  * delta_c = 1e-8 on the obstacle / pair rows (always on, shared with the kernels); on all other rows 0, and 1e-9 only when
    the factorisation of the unperturbed system breaks down (a constant 1e-9 stalls the constraint violation at
    delta_c * |dy| once multipliers of nearly dependent collocation rows reach 1e5, as on the 4-vehicle instance).
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

STATUS = {
    0: "Solve_Succeeded",
    1: "Solved_To_Acceptable_Level",
    -1: "Maximum_Iterations_Exceeded",
    -2: "Restoration_Failed",
    -3: "Error_In_Step_Computation",
    -4: "Invalid_Number_Detected",
    -5: "Search_Direction_Becomes_Too_Small",
    -6: "Infeasible_Problem_Detected",
}


@dataclass
class IpmOptions:
    tol: float = 1e-2
    constr_viol_tol: float = 1e-2
    dual_inf_tol: float = 1.0
    compl_inf_tol: float = 1e-4
    max_iter: int = 3000
    mu_init: float = 0.1
    kappa_eps: float = 10.0
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    tau_min: float = 0.99
    bound_push: float = 1e-2
    bound_frac: float = 1e-2
    kappa_sigma: float = 1e10
    kappa_d: float = 1e-4
    s_max: float = 100.0
    gamma_theta: float = 1e-5
    gamma_phi: float = 1e-8
    eta_phi: float = 1e-8
    delta_ls: float = 1.0
    s_theta: float = 1.1
    s_phi: float = 2.3
    gamma_alpha: float = 0.05
    delta_w_first: float = 1e-4
    delta_w_min: float = 1e-20
    delta_w_max: float = 1e40
    kappa_w_minus: float = 1.0 / 3.0
    kappa_w_plus: float = 8.0
    kappa_w_plus_first: float = 100.0
    delta_c_local: float = 1e-8
    delta_c_global: float = 0.0     # IPOPT: delta_c = 0 unless the KKT matrix is singular ...
    delta_c_singular: float = 1e-9  # ... then a tiny delta_c on every row (LICQ fails: stationary vehicle, dependent collocation rows)
    verbose: int = 0


@dataclass
class IpmResult:
    x: np.ndarray
    y: np.ndarray
    zL: np.ndarray
    zU: np.ndarray
    status: int
    iters: int
    obj: float
    cviol: float
    dual_inf: float
    compl: float
    mu: float
    history: list = field(default_factory=list)

    @property
    def return_status(self):
        return STATUS[self.status]


def push_into_bounds(x, xL, xU, k1, k2):
    """IPOPT initial-point projection (bound_push / bound_frac)."""
    x = x.copy()
    hasL, hasU = np.isfinite(xL), np.isfinite(xU)
    both = hasL & hasU
    pL = np.where(hasL, k1 * np.maximum(1.0, np.abs(np.where(hasL, xL, 0.0))), 0.0)
    pU = np.where(hasU, k1 * np.maximum(1.0, np.abs(np.where(hasU, xU, 0.0))), 0.0)
    span = np.where(both, xU - xL, np.inf)
    pL = np.where(both, np.minimum(pL, k2 * span), pL)
    pU = np.where(both, np.minimum(pU, k2 * span), pU)
    x = np.where(hasL, np.maximum(x, xL + pL), x)
    x = np.where(hasU, np.minimum(x, xU - pU), x)
    return x


class KktSolver:
    """Sparse LU of the augmented system + inertia test (is the reduced Hessian positive definite?)."""

    def __init__(self, nlp, opts):
        self.nlp, self.opts = nlp, opts
        self.local = np.zeros(nlp.m, dtype=bool)
        for name in ("r_obs", "r_pair"):
            for r in getattr(nlp, name, []):
                self.local[np.ravel(r)] = True
        self.t_inertia = self.t_solve = 0.0

    def delta_c(self, dc_global):
        return np.where(self.local, self.opts.delta_c_local, dc_global)

    def inertia_ok(self, Hs, J):
        """PD test of Hs + J' J / eps by diagonal-pivot LU (eps = 1e-7): equivalent to inertia (n, m, 0)."""
        M = (Hs + (J.T @ J) * 1e7).tocsc()
        try:
            lu = spla.splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
        except RuntimeError:
            return False
        if not np.array_equal(lu.perm_r, lu.perm_c):
            return False
        return bool(np.all(lu.U.diagonal() > 0))

    def solve(self, W, Sigma, J, rx, rc, delta_w):
        import time

        n, m = self.nlp.n, self.nlp.m
        Hs = (W + sp.diags(Sigma + delta_w)).tocsc()
        t0 = time.perf_counter()
        ok = self.inertia_ok(Hs, J)
        self.t_inertia += time.perf_counter() - t0
        if not ok:
            return None
        t0 = time.perf_counter()
        rhs = np.concatenate([rx, rc])
        sol = None
        for dcg in (self.opts.delta_c_global, self.opts.delta_c_singular):
            K = sp.bmat([[Hs, J.T], [J, -sp.diags(self.delta_c(dcg))]], format="csc")
            try:
                lu = spla.splu(K)
            except RuntimeError:  # exactly singular
                continue
            with np.errstate(all="ignore"):
                sol = lu.solve(rhs)
                for _ in range(3):  # iterative refinement
                    sol = sol + lu.solve(rhs - K @ sol)
                res = np.abs(rhs - K @ sol).max()
            if np.all(np.isfinite(sol)) and res <= 1e-6 * max(1.0, np.abs(rhs).max()):
                break
            sol = None
        self.t_solve += time.perf_counter() - t0
        if sol is None:
            return None
        return sol[:n], sol[n:]


def solve(nlp, x0, opts: IpmOptions = None, kkt_factory=KktSolver) -> IpmResult:
    o = opts or IpmOptions()
    n, m = nlp.n, nlp.m
    xL, xU = nlp.xL, nlp.xU
    hasL, hasU = np.isfinite(xL), np.isfinite(xU)
    onlyL, onlyU = hasL & ~hasU, hasU & ~hasL
    nb = int(hasL.sum() + hasU.sum())
    x = push_into_bounds(np.asarray(x0, dtype=float), xL, xU, o.bound_push, o.bound_frac)
    y = np.zeros(m)
    zL, zU = np.where(hasL, 1.0, 0.0), np.where(hasU, 1.0, 0.0)
    mu = o.mu_init
    kkt = kkt_factory(nlp, o)

    def gaps(x):
        return np.where(hasL, x - xL, 1.0), np.where(hasU, xU - x, 1.0)

    def phi(x, mu):
        gL, gU = gaps(x)
        if np.any(gL <= 0) or np.any(gU <= 0):
            return np.inf
        val = nlp.f(x) - mu * (np.log(gL[hasL]).sum() + np.log(gU[hasU]).sum())
        val += o.kappa_d * mu * (gL[onlyL].sum() + gU[onlyU].sum())
        return val

    theta = lambda c: np.abs(c).sum()
    c = nlp.c(x)
    th0 = theta(c)
    theta_max, theta_min = 1e4 * max(1.0, th0), 1e-4 * max(1.0, th0)
    filt = []
    delta_w_last = 0.0
    tiny_last = False
    force_mu = False
    best = None  # IPOPT StoreAcceptablePoint: best iterate with E0 <= acceptable_tol = 1e-6
    n_acceptable = 0
    hist = []
    status = -1
    it = 0
    while True:
        f = nlp.f(x)
        g = nlp.grad_f(x)
        c = nlp.c(x)
        J = nlp.jac(x)
        gL, gU = gaps(x)
        r_d = g + J.T @ y - zL + zU
        dual_inf = np.abs(r_d).max()
        cviol = np.abs(c).max() if m else 0.0
        compl = lambda mu_: max(np.abs(gL * zL - mu_)[hasL].max(initial=0.0), np.abs(gU * zU - mu_)[hasU].max(initial=0.0))
        s_d = max(o.s_max, (np.abs(y).sum() + zL.sum() + zU.sum()) / max(1, m + nb)) / o.s_max
        s_c = max(o.s_max, (zL.sum() + zU.sum()) / max(1, nb)) / o.s_max
        E = lambda mu_: max(dual_inf / s_d, cviol, compl(mu_) / s_c)
        hist.append(dict(it=it, f=f, theta=theta(c), dual_inf=dual_inf, cviol=cviol, compl=compl(0.0), mu=mu, dw=delta_w_last))
        if o.verbose:
            print("%4d f=%.8e th=%.2e du=%.2e co=%.2e mu=%.1e dw=%.1e" % (it, f, theta(c), dual_inf, compl(0.0), mu, delta_w_last))
        if E(0.0) <= o.tol and dual_inf <= o.dual_inf_tol and cviol <= o.constr_viol_tol and compl(0.0) <= o.compl_inf_tol:
            status = 0
            break
        if E(0.0) <= 1e-6 and cviol <= 1e-2 and compl(0.0) <= 1e-2:
            n_acceptable += 1
            if best is None or E(0.0) < best[0]:
                best = (E(0.0), x.copy(), y.copy(), zL.copy(), zU.copy())
            if n_acceptable >= 15:
                status = 1
                break
        else:
            n_acceptable = 0
        if it >= o.max_iter:
            status = -1
            break
        if not (np.isfinite(f) and np.all(np.isfinite(c))):
            status = -4
            break
        mu_min = min(o.tol, o.compl_inf_tol) / (o.kappa_eps + 1.0)  # IPOPT monotone update floor
        while (E(mu) <= o.kappa_eps * mu or force_mu) and mu > mu_min:
            mu = max(mu_min, min(o.kappa_mu * mu, mu ** o.theta_mu))
            filt = []
            force_mu = False
        force_mu = False
        tau = max(o.tau_min, 1.0 - mu)

        # ---- search direction
        Sigma = np.where(hasL, zL / gL, 0.0) + np.where(hasU, zU / gU, 0.0)
        damp = o.kappa_d * mu * (onlyL.astype(float) - onlyU.astype(float))
        grad_phi = g - np.where(hasL, mu / gL, 0.0) + np.where(hasU, mu / gU, 0.0) + damp
        rx = -(grad_phi + J.T @ y)
        W = nlp.hess(x, y)
        sol = None
        dw = 0.0
        first = True
        while True:
            sol = kkt.solve(W, Sigma, J, rx, -c, dw)
            if sol is not None:
                break
            if first:
                dw = o.delta_w_first if delta_w_last == 0.0 else max(o.delta_w_min, o.kappa_w_minus * delta_w_last)
                first = False
            else:
                dw *= o.kappa_w_plus_first if delta_w_last == 0.0 else o.kappa_w_plus
            if dw > o.delta_w_max:
                break
        if sol is None:
            status = -3
            break
        if dw > 0:
            delta_w_last = dw
        dx, dy = sol
        dzL = np.where(hasL, mu / gL - zL - zL / gL * dx, 0.0)
        dzU = np.where(hasU, mu / gU - zU + zU / gU * dx, 0.0)

        # ---- fraction to the boundary
        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, (-tau * v[neg] / dv[neg]).min(initial=1.0))

        a_pr = min(max_step(gL[hasL], dx[hasL]), max_step(gU[hasU], -dx[hasU]))
        a_du = min(max_step(zL[hasL], dzL[hasL]), max_step(zU[hasU], dzU[hasU]))

        # ---- filter line search
        th, ph = theta(c), phi(x, mu)
        dphi = grad_phi @ dx
        if dphi < 0:
            a_min = min(o.gamma_theta, o.gamma_phi * th / (-dphi))
            if th <= theta_min:
                a_min = min(a_min, o.delta_ls * th ** o.s_theta / (-dphi) ** o.s_phi)
        else:
            a_min = o.gamma_theta
        a_min *= o.gamma_alpha
        alpha = a_pr
        accepted = False
        # IPOPT compares with a machine-precision slack (Compare_le: lhs - rhs <= 10 eps |base|)
        # widened to 1e-10 relative (shared with the CUDA solver, whose structured steps carry ~1e-8 relative noise)
        slack_phi, slack_th = 1e-10 * max(1.0, abs(ph)), 1e-10 * max(1.0, th)
        # tiny-step rule: a step below 10 eps relative size is accepted without line search and forces a mu update
        tiny = np.max(np.abs(dx) / (1.0 + np.abs(x))) < 10 * np.finfo(float).eps
        if tiny:
            if tiny_last and mu <= mu_min:
                # IPOPT: Search_Direction_Becomes_Too_Small unless the iterate passes the acceptable test
                status = 1 if (E(0.0) <= 1e-6 and cviol <= 1e-2 and compl(0.0) <= 1e-2) else -5
                break
            tiny_last = True
            force_mu = True
            xt = x + alpha * dx
            accepted = True
        else:
            tiny_last = False
        while alpha >= a_min and not accepted:
            xt = x + alpha * dx
            ct = nlp.c(xt)
            tht, pht = theta(ct), phi(xt, mu)
            ok = np.isfinite(pht) and tht <= theta_max and all(not (tht - ft > slack_th and pht - fp > slack_phi) for ft, fp in filt)
            if ok:
                switching = th <= theta_min and dphi < 0 and alpha * (-dphi) ** o.s_phi > o.delta_ls * th ** o.s_theta
                if switching:
                    if pht - (ph + o.eta_phi * alpha * dphi) <= slack_phi:
                        accepted = True
                        break
                elif tht - (1 - o.gamma_theta) * th <= slack_th or pht - (ph - o.gamma_phi * th) <= slack_phi:
                    filt.append(((1 - o.gamma_theta) * th, ph - o.gamma_phi * th))
                    accepted = True
                    break
            alpha *= 0.5
        if not accepted:
            status = -2  # becomes Solved_To_Acceptable_Level below when an acceptable point was stored
            break
        x = x + alpha * dx
        y = y + alpha * dy
        zL = zL + a_du * dzL
        zU = zU + a_du * dzU
        gL, gU = gaps(x)
        zL = np.where(hasL, np.clip(zL, mu / (o.kappa_sigma * gL), o.kappa_sigma * mu / gL), 0.0)
        zU = np.where(hasU, np.clip(zU, mu / (o.kappa_sigma * gU), o.kappa_sigma * mu / gU), 0.0)
        hist[-1].update(alpha=alpha, alpha_du=a_du, dw_used=dw)
        it += 1
    if status != 0 and best is not None:
        _, x, y, zL, zU = best  # IPOPT RestoreAcceptablePoint
        status = 1
    # an active elastic variable at a converged point: the penalised problem is solved, the reference problem (hard distance
    # rows) is infeasible there -- IPOPT reports Infeasible_Problem_Detected and Opti raises
    el = x[nlp._elastic_index()] if hasattr(nlp, "_elastic_index") else np.zeros(0)
    el_max = float(el.max(initial=0.0))
    if status >= 0 and el_max > max(o.constr_viol_tol, 1e-8):
        status = -6
    return IpmResult(x, y, zL, zU, status, it, float(nlp.f(x)), float(np.abs(nlp.c(x)).max()), float(dual_inf), float(compl(0.0)), mu, hist)
