"""Data model, geometry and collocation constants (pins from SURVEY.md App. A.2 / A.3 and the reference's own
``test_rectangle_obstacle`` property, confrez/obstacle_types.py:194-209)."""
import numpy as np
import pytest

from conflict_rez_b200.obstacle_types import GeofenceRegion, RectangleObstacle
from conflict_rez_b200.polytope import Polytope
from conflict_rez_b200.pytypes import VehiclePrediction, VehicleState
from conflict_rez_b200.vehicle_types import VehicleBody, VehicleConfig
from conflict_rez_b200.control.compute_sets import compute_obstacles, compute_sets, convert_rl_states
from conflict_rez_b200.control import warmstart
from oracle.collocation import collocation_coefficients, f_ct, f_rk4, radau_points

A_GOLD = np.array(
    [
        [-25.0000000000, -11.0386792412, 3.5830685225, -2.3441715579, 2.2826355002, -5.0000000000],
        [27.7809339441, 8.7559239779, -7.1613807201, 4.1221652462, -3.8786632197, 8.4124242236],
        [-3.6414784980, 2.8919426154, 1.8060777241, -4.4960171258, 3.3931519181, -6.9702561167],
        [1.2525477212, -0.8751863962, 2.3637971761, 0.8567652454, -5.1883409064, 8.7771142042],
        [-0.5920031672, 0.3997052079, -0.8659007803, 2.5183209492, 0.5812330526, -18.2192823111],
        [0.2000000000, -0.1337061638, 0.2743380778, -0.6570627571, 2.8099836553, 13.0000000000],
    ]
)


def test_collocation_constants_match_known_answer():
    tau = np.append(0, radau_points(5))
    assert np.allclose(tau, [0, 0.05710419611451768, 0.27684301363812383, 0.5835904323689168, 0.8602401356562194, 1.0], atol=1e-13)
    A, B, D = collocation_coefficients(5)
    assert np.allclose(A, A_GOLD, atol=5e-9)
    assert np.allclose(B, [0, 0.1437135608, 0.2813560151, 0.3118265230, 0.2231039011, 0.04], atol=5e-10)
    assert np.allclose(D, [0, 0, 0, 0, 0, 1], atol=1e-9)
    assert np.allclose(warmstart.radau_nodes(5), tau, atol=1e-14)


def test_body_and_limits():
    vb = VehicleBody()
    assert np.array_equal(vb.A, [[1, 0], [0, 1], [-1, 0], [0, -1]])
    assert np.allclose(vb.b, [3.3, 0.9, 0.6, 0.9])
    assert (vb.lf, vb.lr, vb.l, vb.w, vb.cf, vb.cr, vb.num_circles) == (3.3, 0.6, pytest.approx(3.9), 1.8, 2.45, -0.2, 4)
    vc = VehicleConfig()
    assert (vc.v_max, vc.a_max, vc.delta_max, vc.w_delta_max) == (2.5, 1.5, 0.85, 1)
    rg = GeofenceRegion()
    assert (rg.x_min, rg.x_max, rg.y_min, rg.y_max) == (2.5, 32.5, 7.5, 27.5)


def test_frozen_messages():
    s = VehicleState()
    s.x.x = 1.0
    with pytest.raises(TypeError):
        s.not_a_field = 1
    p = VehiclePrediction()
    p.x = np.zeros(3)
    q = p.copy()
    q.x[0] = 5
    assert p.x[0] == 0


def test_rectangle_obstacle_property():
    rng = np.random.default_rng(0)
    for _ in range(1000):
        xc, yc, w, h, psi = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(0.1, 3), rng.uniform(0.1, 3), rng.uniform(-np.pi, np.pi)
        r = RectangleObstacle(xc=xc, yc=yc, w=w, h=h, psi=psi)
        for v in r.V:
            assert np.all(r.A @ v <= r.b + 1e-9)
            assert not np.all(r.A @ v <= r.b - 1e-9)


def test_obstacles_match_survey_constants():
    boxes = [(2.85, 14.65, 7.5, 13.75), (17.85, 19.65, 7.5, 13.75), (22.85, 32.15, 7.5, 13.75), (2.85, 14.65, 21.25, 27.5), (17.85, 22.15, 21.25, 27.5), (25.35, 32.15, 21.25, 27.5)]
    for o, (x0, x1, y0, y1) in zip(compute_obstacles(), boxes):
        assert np.allclose(o.V.min(0), [x0, y0]) and np.allclose(o.V.max(0), [x1, y1])
        assert np.allclose(np.linalg.norm(o.A, axis=1), 1.0)
        assert o.b.shape == (4, 1)


def test_polytope_translation_and_membership():
    p = Polytope([[0, 0], [0, 2.5], [2.5, 0], [2.5, 2.5]]) + np.array([5.0, 7.5])
    assert p.contains([6.0, 8.0]) and not p.contains([4.9, 8.0])


def test_strategy_schema_and_sets(strategy_file):
    import pickle

    with open(strategy_file + ".pkl", "rb") as f:
        hist = pickle.load(f)
    assert sorted(hist) == ["vehicle_%d" % i for i in range(4)]
    starts = {"vehicle_0": ((6, 8), (6, 7)), "vehicle_1": ((8, 7), (9, 7)), "vehicle_2": ((6, 5), (6, 4)), "vehicle_3": ((5, 6), (4, 6))}
    goals = {"vehicle_0": ((12, 6), (11, 6)), "vehicle_1": ((6, 3), (6, 4)), "vehicle_2": ((1, 7), (2, 7)), "vehicle_3": ((6, 10), (6, 9))}
    from conflict_rez_b200.control.strategy import wall_cells

    walls = wall_cells()
    for a, steps in hist.items():
        assert (steps[0]["front"], steps[0]["back"]) == starts[a]
        assert (steps[-1]["front"], steps[-1]["back"]) == goals[a]
        for s0, s1 in zip(steps, steps[1:]):
            assert s1["front"] not in walls and s1["back"] not in walls
            assert max(abs(s1["front"][0] - s1["back"][0]), abs(s1["front"][1] - s1["back"][1])) == 1
            assert s1 == s0 or s1["back"] == s0["front"] or s1["front"] == s0["back"]  # stop, forward or backward move
    # no two agents share a cell at the same step
    T = max(len(v) for v in hist.values())
    for t in range(T):
        cells = []
        for a, steps in hist.items():
            if t < len(steps):
                cells += [steps[t]["front"], steps[t]["back"]]
        assert len(cells) == len(set(cells))
    sets = compute_sets(strategy_file)
    st = convert_rl_states(hist["vehicle_0"][0], VehicleBody())
    assert np.allclose([st.x.x, st.x.y, st.e.psi], [16.25, 18.75, np.pi / 2])
    assert sets["vehicle_0"][0]["front"].contains([st.x.x + 2.5 * np.cos(st.e.psi), st.x.y + 2.5 * np.sin(st.e.psi)])


def test_dynamics_known_values():
    z, u = np.array([1.0, 2.0, 0.3, 1.5, 0.2]), np.array([0.4, -0.1])
    assert np.allclose(f_ct(z, u), [1.5 * np.cos(0.3), 1.5 * np.sin(0.3), 1.5 / 2.5 * np.tan(0.2), 0.4, -0.1])
    # RK4 x 4 against a fine Euler integration of the same ODE
    zz, n = z.copy(), 200000
    for _ in range(n):
        zz = zz + 0.1 / n * f_ct(zz, u)
    assert np.allclose(f_rk4(z, u, 0.1), zz, atol=2e-7)


def test_dual_warm_start_is_dual_feasible_and_tight():
    """Closed-form dual_ws: G'mu + R'A'lam = 0, |A'lam| = 1, and the margin equals the rectangle distance."""
    rng = np.random.default_rng(1)
    vb = VehicleBody()
    G, g = np.asarray(vb.A, float), np.asarray(vb.b, float)
    obs = compute_obstacles()
    A = np.stack([o.A for o in obs])
    b = np.stack([o.b.ravel() for o in obs])
    x, y, psi = rng.uniform(15, 20, 50), rng.uniform(15.5, 19.5, 50), rng.uniform(-np.pi, np.pi, 50)
    lam, mu = warmstart.dual_ws_rect(x, y, psi, A, b, G, g)
    assert lam.min() >= 0 and mu.min() >= 0
    for k in range(50):
        R = np.array([[np.cos(psi[k]), -np.sin(psi[k])], [np.sin(psi[k]), np.cos(psi[k])]])
        body = warmstart._body_vertices(x[k], y[k], psi[k], G, g)
        for j in range(6):
            assert np.allclose(G.T @ mu[k, j] + R.T @ A[j].T @ lam[k, j], 0, atol=1e-12)
            assert np.isclose(np.linalg.norm(A[j].T @ lam[k, j]), 1.0)
            margin = -g @ mu[k, j] + (A[j] @ [x[k], y[k]] - b[j]) @ lam[k, j]
            _, _, dist = warmstart.closest_points_convex(body, warmstart._rect_vertices(A[j], b[j]))
            inside = max((A[j] @ v - b[j]).max() for v in body) < 0 or np.all(A[j] @ body.T <= b[j][:, None], axis=0).any()
            if not inside and margin > 1e-9:
                assert np.isclose(margin, dist, atol=1e-9)
