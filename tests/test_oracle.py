"""The oracle itself: derivative checks of the restated NLP, KKT quality of its solutions, golden fixtures."""
import numpy as np
import pytest

from cases import check_solution_properties, load_golden
from conftest import make_case


@pytest.fixture(scope="module")
def small_joint(strategy_file):
    from oracle.nlp import CollocationNLP

    prob, guess = make_case(strategy_file, ("vehicle_1", "vehicle_2"))
    nlp = CollocationNLP(prob)
    return prob, guess, nlp


def test_problem_dimensions_match_survey_formulas(small_joint):
    """n, eq, ineq of the reference formulation (the oracle adds one slack per inequality and one
    elastic variable per distance row)."""
    prob, _, nlp = small_joint
    O, nodes, N, S = prob.O, prob.nodes, prob.N, prob.n_sets
    n_ref = sum((7 + 8 * O) * m for m in nodes) + 1 + 10 * int(nodes.min())
    ineq = sum(O * m + 8 * (s - 1) for m, s in zip(nodes, S)) + 2 * int(nodes.min())
    elastic = sum(O * m for m in nodes) + int(nodes.min())
    assert nlp.n == n_ref + ineq + elastic
    eq_ref = sum(5 * m + 7 * (n - 1) + 3 * O * m + 7 + 5 for m, n in zip(nodes, N)) + 4 * int(nodes.min())
    assert nlp.m == eq_ref + ineq


def test_jacobian_and_hessian_against_finite_differences(small_joint):
    _, guess, nlp = small_joint
    rng = np.random.default_rng(0)
    x = nlp.init_slacks(nlp.pack(guess)) + 0.01 * rng.standard_normal(nlp.n)
    y = rng.standard_normal(nlp.m)
    J, H = nlp.jac(x), nlp.hess(x, y, clip=False)
    assert abs(H - H.T).max() < 1e-12
    gl = lambda xx: nlp.grad_f(xx) + nlp.jac(xx).T @ y
    for _ in range(3):
        d = rng.standard_normal(nlp.n)
        h = 1e-6
        assert np.abs((nlp.c(x + h * d) - nlp.c(x - h * d)) / (2 * h) - J @ d).max() <= 1e-6 * np.abs(J @ d).max()
        assert abs((nlp.f(x + h * d) - nlp.f(x - h * d)) / (2 * h) - nlp.grad_f(x) @ d) <= 1e-6 * abs(nlp.grad_f(x) @ d)
        assert np.abs((gl(x + h * d) - gl(x - h * d)) / (2 * h) - H @ d).max() <= 1e-5 * np.abs(H @ d).max()


@pytest.mark.parametrize("name", ["single_vehicle_1", "single_vehicle_2", "single_vehicle_2_free_heading", "joint_vehicle_1_2"])
def test_golden_solutions_are_kkt_points_of_the_reference_problem(name):
    """Golden vectors: the stored oracle solutions satisfy the *reference's* problem statement (dynamics, tubes, terminal
    conditions, clearance by plain geometry); the elastic variables vanished, so the exact penalty is exact."""
    prob, _, gold = load_golden(name)
    assert gold["status"] == 0 and gold["cviol"] <= 1e-8 and gold["dual_inf"] <= 1e-6
    worst = check_solution_properties(prob, gold["z"][None], np.array([gold["dt"]]))
    assert worst["collocation"] <= 1e-7 and worst["continuity"] <= 1e-9 and worst["init"] <= 1e-9 and worst["terminal"] <= 1e-9
    assert worst["tube"] <= 1e-7 and worst["bounds"] <= 1e-9
    assert worst["obstacle_clearance"] >= prob.dmin - 1e-6
    if prob.V > 1:
        assert worst["vehicle_clearance"] >= prob.dmin - 1e-6


def test_oracle_reproduces_golden_and_matches_scipy_on_a_tiny_case(strategy_file):
    """(a) the oracle IPM reproduces its committed golden objective; (b) an independent solver (scipy SLSQP) started at
    the oracle solution of a tiny instance (2 sets, 1 obstacle) does not find a better point."""
    from scipy.optimize import minimize

    from oracle import ipm
    from oracle.nlp import CollocationNLP

    prob, guess, gold = load_golden("single_vehicle_2")
    nlp = CollocationNLP(prob)
    res = ipm.solve(nlp, nlp.init_slacks(nlp.pack(guess)), ipm.IpmOptions(tol=1e-8, constr_viol_tol=1e-8, max_iter=300))
    assert res.status == 0 and abs(res.obj - gold["obj"]) <= 1e-9 * abs(gold["obj"])
    assert np.abs(nlp.unpack(res.x)["z"] - gold["z"]).max() <= 1e-7

    tiny, tguess = make_case(strategy_file, ("vehicle_2",))
    tiny.n_sets = np.array([2])
    tiny.obs_A, tiny.obs_b = tiny.obs_A[:1], tiny.obs_b[:1]
    tiny.n_per_set = 2
    M = int(tiny.nodes[0])
    from conflict_rez_b200.problem import CollocationGuess

    idx = np.linspace(0, 29, M).astype(int)
    tg = CollocationGuess(tguess.z[:, idx], tguess.lam[:, idx][:, :, :1], tguess.mu[:, idx][:, :, :1], np.float64(1.0))
    tiny.final_heading = np.array([np.nan])
    tn = CollocationNLP(tiny)
    r = ipm.solve(tn, tn.init_slacks(tn.pack(tg)), ipm.IpmOptions(tol=1e-8, constr_viol_tol=1e-8, max_iter=500))
    assert r.status in (0, 1)
    cons = [{"type": "eq", "fun": tn.c, "jac": lambda x: tn.jac(x).toarray()}]
    bounds = [(None if not np.isfinite(lo) else lo, None if not np.isfinite(hi) else hi) for lo, hi in zip(tn.xL, tn.xU)]
    s = minimize(tn.f, r.x, jac=tn.grad_f, constraints=cons, bounds=bounds, method="SLSQP", options={"maxiter": 200, "ftol": 1e-12})
    assert s.fun >= r.obj - 1e-5 * abs(r.obj) or np.abs(tn.c(s.x)).max() > 1e-6


def test_casadi_ipopt_pins_the_oracle_when_available():
    """on a machine that has CasADi/IPOPT the literal Opti statement of the reference problem
    (oracle/casadi_ref.py) must reproduce the oracle's golden solutions.  Skipped (and the parity stays "unpinned at the
    CasADi/IPOPT boundary") where the wheels are missing -- as in the build image."""
    from oracle import casadi_ref

    if not casadi_ref.available():
        pytest.skip("casadi is not installed here: the oracle is the restated port (parity unpinned at the CasADi/IPOPT boundary)")
    for name in ("single_vehicle_2", "joint_vehicle_1_2"):
        prob, guess, gold = load_golden(name)
        ref = casadi_ref.solve(prob, guess, tol=1e-8)
        assert ref["return_status"] in ("Solve_Succeeded", "Solved_To_Acceptable_Level")
        assert abs(ref["obj"] - gold["obj"]) <= 1e-6 * abs(gold["obj"])
        assert np.abs(ref["z"] - gold["z"]).max() <= 1e-4 and abs(ref["dt"] - gold["dt"]) <= 1e-6
        print(name, "IPOPT iterations", ref["iters"], "oracle", int(gold["iters"]), "linear solver", ref["linear_solver"])
