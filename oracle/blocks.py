"""ORACLE (test infrastructure, never imported by the product path).

Per-block residuals of the OBCA NLPs written once as sympy expressions that transcribe the
reference constraint by constraint; Jacobians and multiplier-contracted Hessians are
derived symbolically, so the oracle shares no hand-written derivative with the CUDA kernels.

Blocks (reference lines in each builder):
  col_k   collocation equations at node k of an interval      vehicle.py:487-509
  cost    running cost  B_k (a^2 + v^2 w^2 + delta^2) dt       vehicle.py:512-521
  obs     obstacle OBCA triple                                  vehicle.py:524-541
  tube    strategy tube sets at a set transition               vehicle.py:570-584
  tubeF   tube set on the end state zF = sum_j D_j z_j         vehicle.py:587-617
  pair    vehicle-vehicle OBCA block                            multi_vehicle_planner.py:419-451

Inequalities are turned into equalities with a non-negative slack (g(x) - s = 0, s >= 0), the
form IPOPT itself uses.
"""
import sympy as sp
import numpy as np

_CACHE = {}


class Block:
    """Lambdified residual ``c``, Jacobian non-zeros and lower-triangular Hessian non-zeros of one block type."""

    def __init__(self, name, exprs, loc, par):
        self.name = name
        self.nloc, self.npar, self.nrow = len(loc), len(par), len(exprs)
        exprs = [sp.sympify(e) for e in exprs]
        J = sp.Matrix(exprs).jacobian(loc)
        self.jac_pat = [(r, c) for r in range(self.nrow) for c in range(self.nloc) if J[r, c] != 0]
        ys = sp.symbols("y0:%d" % self.nrow)
        lag = sum(y * e for y, e in zip(ys, exprs))
        H = sp.hessian(lag, loc)
        self.hes_pat = [(r, c) for r in range(self.nloc) for c in range(r + 1) if H[r, c] != 0]
        args = list(loc) + list(par)
        self._c = sp.lambdify(args, exprs, modules="numpy", cse=True)
        self._j = sp.lambdify(args, [J[r, c] for r, c in self.jac_pat], modules="numpy", cse=True)
        self._h = sp.lambdify(args + list(ys), [H[r, c] for r, c in self.hes_pat], modules="numpy", cse=True)

    @staticmethod
    def _stack(vals, n):
        return np.stack([np.broadcast_to(np.asarray(v, dtype=float), (n,)) for v in vals], axis=0) if vals else np.zeros((0, n))

    def c(self, X, P):
        """X: (nloc, n) locals, P: (npar, n) params -> (nrow, n)."""
        return self._stack(self._c(*X, *P), X.shape[1])

    def jac(self, X, P):
        return self._stack(self._j(*X, *P), X.shape[1])

    def hes(self, X, P, Y):
        return self._stack(self._h(*X, *P, *Y), X.shape[1])


def _rot(psi):
    return sp.Matrix([[sp.cos(psi), -sp.sin(psi)], [sp.sin(psi), sp.cos(psi)]])


def _f_ct(z, u, wb):
    x, y, psi, v, delta = z
    a, w = u
    return sp.Matrix([v * sp.cos(psi), v * sp.sin(psi), v / wb * sp.tan(delta), a, w])


def col_block(k, K=5):
    """Collocation at node k: sum_j A[j,k] z_j / dt - f(z_k, u_k) = 0  (5 rows).

    locals: z_0..z_K (5 each, node-major), u_k (2), dt;  params: A[0..K, k], wb.
    """
    key = ("col", k, K)
    if key not in _CACHE:
        Z = sp.Matrix(K + 1, 5, sp.symbols("z0:%d" % (5 * (K + 1))))
        u = sp.symbols("ua uw")
        dt = sp.Symbol("dt")
        A = sp.symbols("A0:%d" % (K + 1))
        wb = sp.Symbol("wb")
        poly = sum((A[j] * Z[j, :].T for j in range(K + 1)), sp.zeros(5, 1)) / dt
        res = poly - _f_ct(list(Z[k, :]), u, wb)
        _CACHE[key] = Block("col%d" % k, list(res), list(Z) + list(u) + [dt], list(A) + [wb])
    return _CACHE[key]


def cost_block():
    """Running cost term B_k (a^2 + v^2 w^2 + delta^2) dt; locals: v, delta, a, w, dt; params: B_k."""
    if "cost" not in _CACHE:
        v, d, a, w, dt, B = sp.symbols("v d a w dt B")
        _CACHE["cost"] = Block("cost", [B * (a ** 2 + v ** 2 * w ** 2 + d ** 2) * dt], [v, d, a, w, dt], [B])
    return _CACHE["cost"]


def obs_block(h=4):
    """Obstacle triple at one node.

    locals: x, y, psi, lam[h], mu[4], sd, el;  params: A (h x 2 row-major), b (h), G (4 x 2), g (4), dmin.
      c1: -g'mu + (A t - b)'lam - dmin - sd + el = 0     (reference: >= dmin; el >= 0 is the elastic
          variable of the exact l1 penalty, see DESIGN.md "elastic inequality rows")
      c2:  G'mu + R(psi)' A' lam = 0               (2 rows)
      c3:  |A' lam|^2 - 1 = 0
    """
    key = ("obs", h)
    if key not in _CACHE:
        x, y, psi = sp.symbols("x y psi")
        lam = sp.Matrix(sp.symbols("l0:%d" % h))
        mu = sp.Matrix(sp.symbols("m0:4"))
        sd, el = sp.symbols("sd el")
        A = sp.Matrix(h, 2, sp.symbols("A0:%d" % (2 * h)))
        b = sp.Matrix(sp.symbols("b0:%d" % h))
        G = sp.Matrix(4, 2, sp.symbols("G0:8"))
        g = sp.Matrix(sp.symbols("g0:4"))
        dmin = sp.Symbol("dmin")
        t = sp.Matrix([x, y])
        c1 = (-g.T * mu)[0] + ((A * t - b).T * lam)[0] - dmin - sd + el
        c2 = G.T * mu + _rot(psi).T * A.T * lam
        Atl = A.T * lam
        c3 = (Atl.T * Atl)[0] - 1
        _CACHE[key] = Block(
            "obs", [c1, c2[0], c2[1], c3], [x, y, psi] + list(lam) + list(mu) + [sd, el], list(A) + list(b) + list(G) + list(g) + [dmin]
        )
    return _CACHE[key]


def _tube_rows(x, y, psi, Ab, bb, Af, bf, wb, ts):
    back = sp.Matrix([x, y])
    front = sp.Matrix([x + wb * sp.cos(psi), y + wb * sp.sin(psi)])
    rb = bb - Ab * back
    rf = bf - Af * front
    return [rb[i] - ts[i] for i in range(4)] + [rf[i] - ts[4 + i] for i in range(4)]


def _tube_params():
    Ab = sp.Matrix(4, 2, sp.symbols("Ab0:8"))
    bb = sp.Matrix(sp.symbols("bb0:4"))
    Af = sp.Matrix(4, 2, sp.symbols("Af0:8"))
    bf = sp.Matrix(sp.symbols("bf0:4"))
    wb = sp.Symbol("wb")
    return Ab, bb, Af, bf, wb


def tube_block():
    """Tube at a set transition: (b_back - shrink) - A_back (x,y) - s = 0, same for the front axle point (8 rows).

    locals: x, y, psi, ts[8]; params: A_back(8), b_back(4, shrink already subtracted), A_front(8), b_front(4), wb.
    """
    if "tube" not in _CACHE:
        x, y, psi = sp.symbols("x y psi")
        ts = sp.symbols("ts0:8")
        Ab, bb, Af, bf, wb = _tube_params()
        rows = _tube_rows(x, y, psi, Ab, bb, Af, bf, wb, ts)
        _CACHE["tube"] = Block("tube", rows, [x, y, psi] + list(ts), list(Ab) + list(bb) + list(Af) + list(bf) + [wb])
    return _CACHE["tube"]


def tubeF_block(K=5):
    """Tube on the end state zF = sum_j D_j z_{N-1,j}; locals: (x,y,psi)_j for j=0..K, ts[8]; params: tube..., D[K+1]."""
    key = ("tubeF", K)
    if key not in _CACHE:
        P = sp.Matrix(K + 1, 3, sp.symbols("p0:%d" % (3 * (K + 1))))
        ts = sp.symbols("ts0:8")
        Ab, bb, Af, bf, wb = _tube_params()
        D = sp.symbols("D0:%d" % (K + 1))
        zF = [sum(D[j] * P[j, c] for j in range(K + 1)) for c in range(3)]
        rows = _tube_rows(zF[0], zF[1], zF[2], Ab, bb, Af, bf, wb, ts)
        _CACHE[key] = Block("tubeF", rows, list(P) + list(ts), list(Ab) + list(bb) + list(Af) + list(bf) + [wb] + list(D))
    return _CACHE[key]


def pair_block(other_is_param=False):
    """Vehicle-vehicle block at one node (multi_vehicle_planner.py:419-451; vehicle_follower.py:322-352).

    locals: (x,y,psi)_a, [(x,y,psi)_b unless parameters], lam[4], mu[4], s[2], sd, sn, el
    params: G(8), g(4), dmin [, (x,y,psi)_b]
      A_i = G R(-psi_i),  b_i = G R(-psi_i) t_i + g
      d : -b_a'lam - b_b'mu - dmin - sd + el = 0   (el >= 0 elastic, as in obs_block)
      e1: A_a'lam + s = 0 (2);  e2: A_b'mu - s = 0 (2)
      n : 1 - s's - sn = 0
    """
    key = ("pair", other_is_param)
    if key not in _CACHE:
        pa = sp.symbols("xa ya pa")
        pb = sp.symbols("xb yb pb")
        lam = sp.Matrix(sp.symbols("l0:4"))
        mu = sp.Matrix(sp.symbols("m0:4"))
        s = sp.Matrix(sp.symbols("s0:2"))
        sd, sn, el = sp.symbols("sd sn el")
        G = sp.Matrix(4, 2, sp.symbols("G0:8"))
        g = sp.Matrix(sp.symbols("g0:4"))
        dmin = sp.Symbol("dmin")

        def Ab(p):
            A = G * _rot(-p[2])
            return A, A * sp.Matrix([p[0], p[1]]) + g

        Aa, ba = Ab(pa)
        Abb, bb = Ab(pb)
        d = -(ba.T * lam)[0] - (bb.T * mu)[0] - dmin - sd + el
        e1 = Aa.T * lam + s
        e2 = Abb.T * mu - s
        n = 1 - (s.T * s)[0] - sn
        loc = list(pa) + ([] if other_is_param else list(pb)) + list(lam) + list(mu) + list(s) + [sd, sn, el]
        par = list(G) + list(g) + [dmin] + (list(pb) if other_is_param else [])
        _CACHE[key] = Block("pair", [d, e1[0], e1[1], e2[0], e2[1], n], loc, par)
    return _CACHE[key]
