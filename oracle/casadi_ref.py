"""ORACLE (test infrastructure, never imported by the product path).

Optional true-reference shim: when ``import casadi`` works on the machine, the collocation OBCA
problem of a ``CollocationProblem`` is stated through ``casadi.Opti`` the way the reference states it -- the same variables,
the same ``subject_to`` list, the same IPOPT options -- and solved by the real IPOPT, so the restated oracle
(oracle/nlp.py + oracle/ipm.py) can be pinned against the reference's own numerical stack:

    single vehicle   confrez/control/vehicle.py:377-658   (setup_single_final_problem / solve_single_final_problem)
    joint            confrez/control/multi_vehicle_planner.py:365-465 (solve_final_problem_obca)

CasADi / IPOPT / HSL are NOT installed in the build image (no network), so this module has not been executed there; every
report that uses it must say which oracle ran (``available()``).  ``linear_solver`` falls back from ``ma97`` (needs a
user-supplied HSL library) to ``mumps`` (bundled with the CasADi wheels).
"""
from itertools import combinations

import numpy as np


def available() -> bool:
    try:
        import casadi  # noqa: F401

        return True
    except Exception:
        return False


def solve(prob, guess, tol=1e-2, max_iter=3000, linear_solver="ma97", verbose=0):
    """Literal Opti statement of the (joint) collocation OBCA problem; returns a dict like ``oracle.nlp.unpack`` plus the
    IPOPT statistics.  ``prob`` is one instance (``prob.batch is None``), ``guess`` its ``CollocationGuess``."""
    import casadi as ca

    from oracle.collocation import collocation_coefficients

    assert prob.batch is None
    K, V, O = prob.K, prob.V, prob.O
    A, B, D = collocation_coefficients(K)
    opti = ca.Opti()
    dt = opti.variable()                                                     # multi_vehicle_planner.py:366-367 (shared time scale)
    opti.set_initial(dt, float(guess.dt))
    G, g = prob.body_G, prob.body_g
    J = 0
    X, LAM, MU = [], [], []
    for a in range(V):
        N = int(prob.N[a])
        x, y, psi, v, de = [opti.variable(N, K + 1) for _ in range(5)]       # vehicle.py:402-409
        ua, uw = opti.variable(N, K + 1), opti.variable(N, K + 1)
        l = [[opti.variable(4 * O) for _ in range(K + 1)] for _ in range(N)]  # vehicle.py:411-416
        m = [[opti.variable(4 * O) for _ in range(K + 1)] for _ in range(N)]
        X.append((x, y, psi, v, de, ua, uw)), LAM.append(l), MU.append(m)
        z0 = guess.z[a]
        for c, var in enumerate((x, y, psi, v, de, ua, uw)):
            opti.set_initial(var, z0[: N * (K + 1), c].reshape(N, K + 1))    # vehicle.py:629-636
        opti.subject_to(x[0, 0] == prob.init_pose[a, 0])                     # vehicle.py:424-434
        opti.subject_to(y[0, 0] == prob.init_pose[a, 1])
        opti.subject_to(psi[0, 0] == prob.init_pose[a, 2])
        for var in (v, de, ua, uw):
            opti.subject_to(var[0, 0] == 0)
        for i in range(N):
            for k in range(K + 1):
                n = i * (K + 1) + k
                opti.subject_to(opti.bounded(prob.region[0], x[i, k], prob.region[1]))   # vehicle.py:439-478
                opti.subject_to(opti.bounded(prob.region[2], y[i, k], prob.region[3]))
                opti.subject_to(opti.bounded(prob.limits[0], v[i, k], prob.limits[1]))
                opti.subject_to(opti.bounded(prob.limits[2], de[i, k], prob.limits[3]))
                opti.subject_to(opti.bounded(prob.limits[4], ua[i, k], prob.limits[5]))
                opti.subject_to(opti.bounded(prob.limits[6], uw[i, k], prob.limits[7]))
                opti.subject_to(l[i][k] >= 0)                                # vehicle.py:481-485
                opti.subject_to(m[i][k] >= 0)
                opti.set_initial(l[i][k], guess.lam[a, n].ravel())
                opti.set_initial(m[i][k], guess.mu[a, n].ravel())
                state = ca.vertcat(x[i, k], y[i, k], psi[i, k], v[i, k], de[i, k])
                f = ca.vertcat(v[i, k] * ca.cos(psi[i, k]), v[i, k] * ca.sin(psi[i, k]), v[i, k] / prob.wb * ca.tan(de[i, k]), ua[i, k], uw[i, k])
                poly = 0
                for j in range(K + 1):                                       # vehicle.py:499-509
                    poly += A[j, k] * ca.vertcat(x[i, j], y[i, j], psi[i, j], v[i, j], de[i, j]) / dt
                opti.subject_to(poly == f)
                J += B[k] * (ua[i, k] ** 2 + v[i, k] ** 2 * uw[i, k] ** 2 + de[i, k] ** 2) * dt   # vehicle.py:512-521
                t = ca.vertcat(x[i, k], y[i, k])
                R = ca.vertcat(ca.horzcat(ca.cos(psi[i, k]), -ca.sin(psi[i, k])), ca.horzcat(ca.sin(psi[i, k]), ca.cos(psi[i, k])))
                for j in range(O):                                           # vehicle.py:524-541
                    lj, mj = l[i][k][4 * j : 4 * (j + 1)], m[i][k][4 * j : 4 * (j + 1)]
                    Aj, bj = ca.DM(prob.obs_A[j]), ca.DM(prob.obs_b[j])
                    opti.subject_to(ca.dot(-ca.DM(g), mj) + ca.dot(Aj @ t - bj, lj) >= prob.dmin)
                    opti.subject_to(ca.DM(G).T @ mj + R.T @ Aj.T @ lj == np.zeros(2))
                    opti.subject_to(ca.dot(Aj.T @ lj, Aj.T @ lj) == 1)
                _ = state
            if i >= 1:                                                       # vehicle.py:544-584
                zprev, uprev = 0, 0
                for j in range(K + 1):
                    zprev += D[j] * ca.vertcat(x[i - 1, j], y[i - 1, j], psi[i - 1, j], v[i - 1, j], de[i - 1, j])
                    uprev += D[j] * ca.vertcat(ua[i - 1, j], uw[i - 1, j])
                opti.subject_to(zprev == ca.vertcat(x[i, 0], y[i, 0], psi[i, 0], v[i, 0], de[i, 0]))
                opti.subject_to(uprev == ca.vertcat(ua[i, 0], uw[i, 0]))
                q, r = divmod(i, prob.n_per_set)
                if r == 0:
                    _tube(opti, ca, prob, a, q, x[i, 0], y[i, 0], psi[i, 0])
        zF, uF = 0, 0                                                        # vehicle.py:586-626
        for j in range(K + 1):
            zF += D[j] * ca.vertcat(x[N - 1, j], y[N - 1, j], psi[N - 1, j], v[N - 1, j], de[N - 1, j])
            uF += D[j] * ca.vertcat(ua[N - 1, j], uw[N - 1, j])
        _tube(opti, ca, prob, a, int(prob.n_sets[a]) - 1, zF[0], zF[1], zF[2])
        if np.isfinite(prob.final_heading[a]):
            opti.subject_to(zF[2] == float(prob.final_heading[a]))
        opti.subject_to(zF[3] == 0), opti.subject_to(zF[4] == 0)
        opti.subject_to(uF[0] == 0), opti.subject_to(uF[1] == 0)
        J += (N * dt) ** 2                                                   # vehicle.py:638
    PL, PM, PS = [], [], []
    for q, (a, b) in enumerate(combinations(range(V), 2)):                   # multi_vehicle_planner.py:388-451
        Nmin = int(min(prob.N[a], prob.N[b]))
        pl = [[opti.variable(4) for _ in range(K + 1)] for _ in range(Nmin)]
        pm = [[opti.variable(4) for _ in range(K + 1)] for _ in range(Nmin)]
        ps = [[opti.variable(2) for _ in range(K + 1)] for _ in range(Nmin)]
        PL.append(pl), PM.append(pm), PS.append(ps)
        for i in range(Nmin):
            for k in range(K + 1):
                n = i * (K + 1) + k
                lik, mik, sik = pl[i][k], pm[i][k], ps[i][k]
                opti.subject_to(lik >= 0), opti.subject_to(mik >= 0)
                opti.set_initial(lik, guess.pair_lam[q, n]), opti.set_initial(mik, guess.pair_mu[q, n]), opti.set_initial(sik, guess.pair_s[q, n])

                def body(xv, yv, pv):
                    Rm = ca.vertcat(ca.horzcat(ca.cos(-pv), -ca.sin(-pv)), ca.horzcat(ca.sin(-pv), ca.cos(-pv)))
                    Av = ca.DM(G) @ Rm
                    return Av, Av @ ca.vertcat(xv, yv) + ca.DM(g)

                Aa, ba = body(X[a][0][i, k], X[a][1][i, k], X[a][2][i, k])
                Ab, bb = body(X[b][0][i, k], X[b][1][i, k], X[b][2][i, k])
                opti.subject_to(-ca.dot(ba, lik) - ca.dot(bb, mik) >= prob.dmin)
                opti.subject_to(Aa.T @ lik + sik == np.zeros(2))
                opti.subject_to(Ab.T @ mik - sik == np.zeros(2))
                opti.subject_to(ca.dot(sik, sik) <= 1)
    opti.minimize(J)
    s_opts = {"print_level": verbose, "tol": tol, "constr_viol_tol": tol, "max_iter": max_iter, "linear_solver": linear_solver}
    used = linear_solver
    try:
        opti.solver("ipopt", {"expand": True}, s_opts)                       # vehicle.py:648-657
        sol = opti.solve()
    except RuntimeError as e:
        if linear_solver == "ma97" and ("ma97" in str(e).lower() or "hsl" in str(e).lower() or "linear solver" in str(e).lower()):
            s_opts["linear_solver"] = used = "mumps"
            opti.solver("ipopt", {"expand": True}, s_opts)
            sol = opti.solve()
        else:
            raise
    Mmax = int(prob.nodes.max())
    out = {"z": np.zeros((V, Mmax, 7)), "lam": np.zeros((V, Mmax, O, 4)), "mu": np.zeros((V, Mmax, O, 4)), "dt": float(sol.value(dt))}
    for a in range(V):
        N = int(prob.N[a])
        for c, var in enumerate(X[a]):
            out["z"][a, : N * (K + 1), c] = np.asarray(sol.value(var)).reshape(-1)
        for i in range(N):
            for k in range(K + 1):
                out["lam"][a, i * (K + 1) + k] = np.asarray(sol.value(LAM[a][i][k])).reshape(O, 4)
                out["mu"][a, i * (K + 1) + k] = np.asarray(sol.value(MU[a][i][k])).reshape(O, 4)
    st = sol.stats()
    out.update(obj=float(sol.value(J)), return_status=st["return_status"], iters=int(st["iter_count"]), t_wall=float(st.get("t_wall_total", np.nan)),
               linear_solver=used)
    return out


def _tube(opti, ca, prob, a, q, x, y, psi):
    """rl_tube[q] constraints on the rear-axle point and the front point (vehicle.py:570-584, 605-617)."""
    back = ca.vertcat(x, y)
    front = ca.vertcat(x + prob.wb * ca.cos(psi), y + prob.wb * ca.sin(psi))
    opti.subject_to(ca.DM(prob.tube_A[a, q, 0]) @ back <= ca.DM(prob.tube_b[a, q, 0]) - prob.shrink_tube)
    opti.subject_to(ca.DM(prob.tube_A[a, q, 1]) @ front <= ca.DM(prob.tube_b[a, q, 1]) - prob.shrink_tube)
