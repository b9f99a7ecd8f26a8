"""State warm start ``Vehicle.state_ws`` (confrez/control/vehicle.py:99-231): the Euler-discretised tube-following NLP.

* the oracle restatement (oracle/state_ws.py) has consistent derivatives (finite differences) and its solutions satisfy the
  reference problem statement checked by plain arithmetic (Euler recursion, tube sets, boundary rows);
* the device solver (OBCA_MODE_STATE_WS, through the C ABI) reproduces the oracle: residuals / Lagrangian gradient / Newton step at
  a perturbed point, full solves (objective, trajectory, iteration count) for every agent, with and without ``bounded_input``,
  batched over initial offsets;
* ``Vehicle.state_ws`` returns what the reference returns (uniform grid, padded inputs).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from conflict_rez_b200.control import warmstart
from conflict_rez_b200.control.compute_sets import compute_initial_states, compute_sets, interp_along_sets
from conflict_rez_b200.solver import ObcaStateWsSolver, SolveOptions, StateWsProblem
from conflict_rez_b200.vehicle_types import VehicleBody
from layout_map import DeviceLayout
from oracle import ipm
from oracle.state_ws import EulerWsNLP

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
HEADINGS = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": None}  # one agent without a final heading
N, DT, SHRINK = 30, 0.1, 0.5


@pytest.fixture(params=BACKENDS)
def backend(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_lib"), "cpu"
    return request.getfixturevalue("cuda_lib"), "cuda:0"


def _case(fn, agent, bounded=False, offsets=None):
    vb = VehicleBody()
    tube = compute_sets(fn)[agent]
    init = compute_initial_states(fn, vb)[agent]
    S = len(tube)
    A, b = np.zeros((S, 2, 4, 2)), np.zeros((S, 2, 4))
    for q, s in enumerate(tube):
        for ib, body in enumerate(("back", "front")):
            A[q, ib], b[q, ib] = s[body].A, np.ravel(s[body].b)
    offsets = np.zeros((1, 3)) if offsets is None else np.asarray(offsets, float)
    B = len(offsets)
    p = StateWsProblem(tube_A=A, tube_b=b, N=N, dt=DT, final_heading=HEADINGS[agent], bounded_input=bounded, shrink_tube=SHRINK, batch=B)
    path = interp_along_sets(fn, vb, N)[agent]
    M = p.nodes
    z0, cur = np.zeros((B, M, 7)), np.zeros((B, 5))
    for i, off in enumerate(offsets):
        pth = path + np.clip(1.0 - np.arange(M) / float(N), 0.0, 1.0)[:, None] * off[None, :]
        kin = warmstart.kinematic_guess(pth, DT, vb.wb, p.limits)
        for c, k in enumerate(("x", "y", "psi", "v", "delta", "a", "w")):
            z0[i, :, c] = kin[k]
        cur[i, :3] = [init.x.x + off[0], init.x.y + off[1], init.e.psi + off[2]]
    nlps = [EulerWsNLP(dict(N=N, dt=DT, wb=vb.wb, tube_A=A, tube_b=b - SHRINK, init=cur[i, :3], heading=HEADINGS[agent], region=p.region, limits=p.limits,
                            bounded_input=bounded)) for i in range(B)]
    return p, cur, z0, nlps


def _oracle_solve(nlp, z0, tol):
    return ipm.solve(nlp, nlp.init_slacks(nlp.pack(z0[:, :5], z0[:-1, 5:])), ipm.IpmOptions(tol=tol, constr_viol_tol=tol, max_iter=500))


def test_oracle_derivatives_and_problem_statement(strategy_file):
    p, cur, z0, (nlp,) = _case(strategy_file, "vehicle_1")
    rng = np.random.default_rng(0)
    x, y = 0.3 * rng.normal(size=nlp.n), rng.normal(size=nlp.m)
    J, H = nlp.jac(x).toarray(), nlp.hess(x, y).toarray()
    eps = 1e-6
    for k in range(0, nlp.n, 3):
        d = np.zeros(nlp.n)
        d[k] = eps
        assert np.abs((nlp.c(x + d) - nlp.c(x - d)) / (2 * eps) - J[:, k]).max() < 1e-7
        gfd = ((nlp.grad_f(x + d) + nlp.jac(x + d).T @ y) - (nlp.grad_f(x - d) + nlp.jac(x - d).T @ y)) / (2 * eps)
        assert np.abs(gfd - H[:, k]).max() < 1e-6
    # the solution against the reference's problem statement, in plain arithmetic (vehicle.py:131-195)
    r = _oracle_solve(nlp, z0[0], 1e-8)
    assert r.status == 0
    z, u = r.x[nlp.iz], r.x[nlp.iu]
    wb = 2.5
    f = np.stack([z[:-1, 3] * np.cos(z[:-1, 2]), z[:-1, 3] * np.sin(z[:-1, 2]), z[:-1, 3] / wb * np.tan(z[:-1, 4]), u[:, 0], u[:, 1]], axis=1)
    assert np.abs(z[1:] - z[:-1] - DT * f).max() <= 1e-8
    assert np.abs(z[0] - cur[0]).max() <= 1e-8 and np.abs(u[0]).max() <= 1e-8 and abs(z[-1, 2] - HEADINGS["vehicle_1"]) <= 1e-8
    assert abs(r.obj - (u ** 2).sum()) <= 1e-12
    for i in range(1, p.n_sets):
        k = N * i
        back, front = z[k, :2], z[k, :2] + wb * np.array([np.cos(z[k, 2]), np.sin(z[k, 2])])
        assert (p.tube_A[i, 0] @ back <= p.tube_b[i, 0] - SHRINK + 1e-7).all() and (p.tube_A[i, 1] @ front <= p.tube_b[i, 1] - SHRINK + 1e-7).all()
    assert (z[:-1, 3] >= p.limits[0] - 1e-9).all() and (z[:-1, 3] <= p.limits[1] + 1e-9).all()


def _maps(L, nlp, M, S):
    D = DeviceLayout(L)
    ix, iy = np.zeros(nlp.n, dtype=int), np.zeros(nlp.m, dtype=int)
    k = np.arange(M)
    for c in range(5):
        ix[nlp.iz[:, c]] = D.Z(0, c, k)
        iy[nlp.r_dyn[:, c]] = D.YCOL(0, c, k[:-1])
    for c in range(2):
        ix[nlp.iu[:, c]] = D.Z(0, 5 + c, k[:-1])
    for q in range(S - 1):
        for r in range(8):
            ix[nlp.its[q, r]], iy[nlp.r_tube[q, r]] = D.TS(0, q, r), D.YTUBE(0, q, r)
    for q in range(7):
        iy[nlp.r_init[q]] = L["oYINIT"] + q
    if len(nlp.r_head):
        iy[nlp.r_head[0]] = L["oYTERM"]
    return ix, iy


def test_residuals_gradient_and_newton_step(backend, strategy_file):
    lib, dev = backend
    p, cur, z0, (nlp,) = _case(strategy_file, "vehicle_1")
    sv = ObcaStateWsSolver(p, SolveOptions(max_iter=0), device=dev, lib=lib)
    sv.solve_ws(cur, z0)  # sets parameters and iterate (no iteration)
    L, M = sv.layout(), p.nodes
    assert L["m_active"] == nlp.m and L["nx"] == nlp.n + 2 + 1  # + the inputs of the last node (cost only, zero at the optimum) + dt slot
    ix, iy = _maps(L, nlp, M, p.n_sets)
    rng = np.random.default_rng(1)
    x0 = nlp.init_slacks(nlp.pack(z0[0, :, :5] + 0.01 * rng.normal(size=(M, 5)), z0[0, :-1, 5:] + 0.1 * rng.normal(size=(M - 1, 2))))
    x0[nlp.its] = np.abs(x0[nlp.its]) + 0.1
    y0 = 0.5 * rng.normal(size=nlp.m)
    xd, yd, zL, zU = sv.debug_get_iterate(0)
    xd[:], yd[:], zL[:], zU[:] = 0, 0, 0, 0
    xd[ix], yd[iy] = x0, y0
    hasL, hasU = np.isfinite(nlp.xL), np.isfinite(nlp.xU)
    zL[ix], zU[ix] = hasL.astype(float), hasU.astype(float)
    sv.debug_set_iterate(0, xd, yd, zL, zU)
    c_d, gl_d, f_d = sv.debug_eval(0)
    c_o, gl_o = nlp.c(x0), nlp.grad_f(x0) + nlp.jac(x0).T @ y0
    assert abs(f_d - nlp.f(x0)) <= 1e-12 * max(1.0, abs(f_d))
    assert np.abs(c_d[iy] - c_o).max() <= 1e-12 and np.abs(gl_d[ix] - gl_o).max() <= 1e-12 * max(1.0, np.abs(gl_o).max())
    mu = 0.1
    dx_d, dy_d, ok = sv.debug_step(0, mu, 0.0)
    assert ok == 1
    gL, gU = x0 - nlp.xL, nlp.xU - x0
    sig, gphi = np.zeros(nlp.n), gl_o.copy()
    sig[hasL] += 1.0 / gL[hasL]
    sig[hasU] += 1.0 / gU[hasU]
    gphi[hasL] -= mu / gL[hasL]
    gphi[hasU] += mu / gU[hasU]
    gphi[hasL & ~hasU] += 1e-4 * mu  # kappa_d
    K = sp.bmat([[nlp.hess(x0, y0) + sp.diags(sig), nlp.jac(x0).T], [nlp.jac(x0), None]], format="csc")
    sol = spl.spsolve(K, -np.concatenate([gphi, c_o]))
    dx_o, dy_o = sol[: nlp.n], sol[nlp.n:]
    # exact elimination of the fixed inputs of stage 0 and of the final-heading row (second right-hand side): sparse-LU accuracy
    assert np.abs(dx_d[ix] - dx_o).max() <= 1e-9 * np.abs(dx_o).max() and np.abs(dy_d[iy] - dy_o).max() <= 1e-9 * np.abs(dy_o).max()
    sv.close()


@pytest.mark.parametrize("agent,bounded", [("vehicle_0", False), ("vehicle_1", False), ("vehicle_1", True), ("vehicle_3", False)])
def test_solution_matches_oracle(backend, strategy_file, agent, bounded):
    lib, dev = backend
    p, cur, z0, (nlp,) = _case(strategy_file, agent, bounded)
    for tol in (1e-2, 1e-8):  # the reference's tolerance (vehicle.py:209-210) and a tight one
        sv = ObcaStateWsSolver(p, SolveOptions(tol=tol, constr_viol_tol=tol, max_iter=500), device=dev, lib=lib)
        res = sv.solve_ws(cur, z0)
        sv.close()
        ref = _oracle_solve(nlp, z0[0], tol)
        print(agent, "bounded" if bounded else "", "tol %.0e: iterations device %d oracle %d" % (tol, res.iters[0], ref.iters))
        assert res.status[0] == 0 and ref.status == 0
        assert abs(res.obj[0] - ref.obj) <= 1e-8 * abs(ref.obj)
        assert np.abs(res.z[0, 0, :, :5] - ref.x[nlp.iz]).max() <= 1e-6 and np.abs(res.z[0, 0, :-1, 5:] - ref.x[nlp.iu]).max() <= 1e-6
        assert np.abs(res.z[0, 0, -1, 5:]).max() <= 1e-8  # the extra inputs of the last node: cost only
        assert abs(int(res.iters[0]) - ref.iters) <= 2
    if bounded:
        assert (np.abs(res.z[0, 0, :, 5]) <= 1.5 + 1e-9).all() and (np.abs(res.z[0, 0, :, 6]) <= 1.0 + 1e-9).all()


def test_batch_of_initial_offsets(backend, strategy_file):
    """One launch, several initial offsets (the bench draws them i.i.d.): each instance equals its own oracle solve."""
    lib, dev = backend
    offs = np.array([[0.0, 0.0, 0.0], [0.1, -0.12, 0.1], [-0.15, 0.05, -0.15]])
    p, cur, z0, nlps = _case(strategy_file, "vehicle_2", offsets=offs)
    sv = ObcaStateWsSolver(p, SolveOptions(tol=1e-8, constr_viol_tol=1e-8, max_iter=500), device=dev, lib=lib)
    res = sv.solve_ws(cur, z0)
    sv.close()
    assert (res.status == 0).all()
    for i, nlp in enumerate(nlps):
        ref = _oracle_solve(nlp, z0[i], 1e-8)
        assert ref.status == 0 and abs(res.obj[i] - ref.obj) <= 1e-8 * abs(ref.obj)
        assert np.abs(res.z[i, 0, :, :5] - ref.x[nlp.iz]).max() <= 1e-6


def test_vehicle_state_ws_returns_the_reference_layout(backend, strategy_file):
    from conflict_rez_b200.control.vehicle import Vehicle
    from conflict_rez_b200.pytypes import VehicleState

    lib, dev = backend
    v = Vehicle(strategy_file, "vehicle_1", {}, device=dev)
    v._lib = lib
    off = VehicleState()
    off.x.x, off.e.psi = 0.1, np.pi / 20  # multi_vehicle_planner.py:641-642
    zu0 = v.state_ws(N=N, dt=DT, init_offset=off, final_heading=3 * np.pi / 2, shrink_tube=SHRINK, spline_ws=True)
    M = N * (v.num_sets - 1) + 1
    assert v.state_ws_result.status[0] == 0
    assert np.allclose(zu0.t, np.linspace(0, (M - 1) * DT, M))
    for k in ("x", "y", "psi", "v", "u_a", "u_steer", "u_steer_dot"):
        assert len(getattr(zu0, k)) == M
    assert zu0.u_a[-1] == zu0.u_a[-2] and zu0.u_steer_dot[-1] == zu0.u_steer_dot[-2]  # vehicle.py:227-229
    assert abs(zu0.x[0] - (v.init_state.x.x + 0.1)) <= 1e-6 and abs(zu0.psi[0] - (v.init_state.e.psi + np.pi / 20)) <= 1e-6
    assert abs(zu0.psi[-1] - 3 * np.pi / 2) <= 1e-2 and zu0.u_a[0] == pytest.approx(0.0, abs=1e-6)
    # Euler consistency of what is returned
    assert np.abs(np.diff(zu0.v) - DT * zu0.u_a[:-1]).max() <= 1e-2
