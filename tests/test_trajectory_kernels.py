"""Trajectory-side kernels (csrc/obca_traj.h) through the C ABI: collocation interpolation (Vehicle.interpolate_states,
confrez/control/vehicle.py:722-829), the MPC reference window (vehicle_follower.py:370-404), the plant step (simulator,
dynamic_model.py:61-93) and the horizon shift (_adv_onestep, vehicle_follower.py:413-426)."""
import numpy as np
import pytest
import torch

from cases import load_golden

from conflict_rez_b200.solver import TrajectoryOps

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def ops(request):
    if request.param == "emu":
        return TrajectoryOps(device="cpu", lib=request.getfixturevalue("emu_lib"))
    return TrajectoryOps(device="cuda:0", lib=request.getfixturevalue("cuda_lib"))


def _dev(ops, a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(ops.device)


def _direct_lagrange(z, dt, N, tau, t):
    """Plain restatement of the reference interpolator for one sample (vehicle.py:738-786)."""
    tgrid = np.linspace(0, N * dt, N + 1)
    i = int(np.searchsorted(tgrid[1:], t, side="right"))
    out = np.zeros(7)
    if i >= N:
        out[:5] = z[N * 6 - 1, :5]
    else:
        rel = (t - tgrid[i]) / dt
        for j in range(6):
            Lj = 1.0
            for k in range(6):
                if k != j:
                    Lj *= (rel - tau[k]) / (tau[j] - tau[k])
            out[:5] += Lj * z[i * 6 + j, :5]
    t_nodes = (np.arange(N)[:, None] + tau[None, :]).ravel() * dt
    m = min(int(np.searchsorted(t_nodes[1:], t, side="right")), N * 6 - 1)
    out[5:] = z[m, 5:]
    return out


def test_interpolation_matches_direct_lagrange_evaluation_and_the_host_planner(ops, strategy_file):
    from conflict_rez_b200.control.warmstart import radau_nodes

    prob, _, gold = load_golden("joint_vehicle_1_2")  # ragged: N = 20 and 25 intervals
    z = gold["z"]
    dt = float(gold["dt"])
    N = [int(n) for n in prob.N]
    tau = radau_nodes(5)
    rng = np.random.default_rng(0)
    T = 97
    # (sample times exactly on an interior interval boundary are left out: the numerically computed Radau node tau_K is
    # 1 + 1e-16, so the node-time array is unsorted by one ulp there and numpy's binary search is not a specification)
    times = np.sort(np.concatenate([rng.uniform(-0.1, max(N) * dt * 1.1, T - 6), [0.0, 0.999 * dt, 3.001 * dt, N[0] * dt * 1.0001, N[1] * dt * 1.0001, 1e3]]))
    out = ops.interpolate(_dev(ops, z[None]), _dev(ops, [dt]), N, _dev(ops, times)).cpu().numpy()[0]
    for a in range(2):
        ref = np.stack([_direct_lagrange(z[a], dt, N[a], tau, t) for t in times])
        assert np.abs(out[a] - ref).max() <= 1e-12
    # node values are reproduced exactly, inputs are the node inputs (piecewise constant)
    t_nodes = (np.arange(N[0])[:, None] + tau[None, :]).ravel() * dt
    on_nodes = ops.interpolate(_dev(ops, z[None]), _dev(ops, [dt]), N, _dev(ops, t_nodes)).cpu().numpy()[0, 0]
    keep = np.concatenate([np.diff(t_nodes) > 1e-12, [True]])  # at a shared interval boundary the later node wins (pw_const)
    assert np.abs(on_nodes[keep, :5] - z[0, : N[0] * 6][keep, :5]).max() <= 1e-12
    # per-vehicle time arrays
    tv = np.stack([times, times[::-1]])[None]
    out2 = ops.interpolate(_dev(ops, z[None]), _dev(ops, [dt]), N, _dev(ops, tv)).cpu().numpy()[0]
    assert np.array_equal(out2[0], out[0]) and np.array_equal(out2[1], out[1][::-1])


def test_reference_window_times_follow_numpy_argmin(ops):
    rng = np.random.default_rng(3)
    B, V, N, dtm = 3, 4, 30, 0.1
    grid = np.zeros((B, V, 3))
    grid[..., 0] = rng.uniform(0, 0.5, (B, V))
    grid[..., 1] = grid[..., 0] + rng.uniform(5, 20, (B, V))
    grid[..., 2] = rng.integers(200, 2500, (B, V))
    clock = rng.uniform(-0.2, 22, (B, V))
    clock[0, 0] = grid[0, 0, 0] + 0.5 * (grid[0, 0, 1] - grid[0, 0, 0]) / (grid[0, 0, 2] - 1)  # a tie between two samples
    out = ops.mpc_ref_times(_dev(ops, grid), _dev(ops, clock), N, dtm).cpu().numpy()
    for b in range(B):
        for v in range(V):
            tref = np.linspace(grid[b, v, 0], grid[b, v, 1], int(grid[b, v, 2]))
            t0 = tref[np.abs(tref - clock[b, v]).argmin()]
            assert np.abs(out[b, v] - (t0 + np.linspace(0, N * dtm, N, endpoint=False))).max() <= 1e-12


def test_plant_step_is_rk4_and_converged(ops):
    from oracle.collocation import f_rk4

    rng = np.random.default_rng(1)
    s = rng.uniform(-1, 1, (16, 5)) * np.array([5, 5, 3, 2.5, 0.8])
    u = rng.uniform(-1, 1, (16, 2)) * np.array([1.5, 1.0])
    four = ops.plant_step(_dev(ops, s), _dev(ops, u), 0.1, 2.5, substeps=4).cpu().numpy()
    for b in range(16):  # the same scheme as kinematic_bicycle_rk (dynamic_model.py:30-58): bit-level agreement up to rounding
        assert np.abs(four[b] - f_rk4(s[b], u[b], 0.1)).max() <= 1e-13
    fine = ops.plant_step(_dev(ops, s), _dev(ops, u), 0.1, 2.5, substeps=100).cpu().numpy()
    finer = ops.plant_step(_dev(ops, s), _dev(ops, u), 0.1, 2.5, substeps=400).cpu().numpy()
    assert np.abs(fine - finer).max() <= 1e-10  # the default plant integrator is converged to the 1e-10 SURVEY.md 8a3 asks for


def test_horizon_shift_is_adv_onestep(ops):
    from conflict_rez_b200.control.vehicle_follower import VehicleFollower

    rng = np.random.default_rng(2)
    a = rng.standard_normal((3, 30, 6, 4))
    out = ops.shift_horizon(_dev(ops, a)).cpu().numpy()
    for b in range(3):
        assert np.array_equal(out[b].reshape(30, -1), VehicleFollower._adv_onestep(a[b].reshape(30, -1)))
    v = rng.standard_normal((2, 30))
    assert np.array_equal(ops.shift_horizon(_dev(ops, v)).cpu().numpy()[1], VehicleFollower._adv_onestep(v[1]))


def test_warm_start_interpolation_is_scipy_interp1d(ops):
    """obca_interp_ws == Vehicle.interp_ws_for_collocation (vehicle.py:298-358): scipy's linear interp1d of every signal onto
    t_interp = (i + tau_k) / N * t[-1]; uniform and non-uniform sample grids, end points included."""
    from scipy.interpolate import interp1d

    from conflict_rez_b200.control.warmstart import interp_ws_for_collocation, radau_nodes

    rng = np.random.default_rng(5)
    dev = ops.device
    for T, N, t in ((271, 45, np.linspace(0, 27.0, 271)), (40, 7, np.sort(np.concatenate([[0.0], rng.uniform(0.01, 3.9, 38), [4.0]])))):
        sig = rng.normal(size=(3, T, 9))
        out = ops.interp_ws(torch.tensor(sig, device=dev), torch.tensor(t, device=dev), N).cpu().numpy()
        ti = (np.arange(N)[:, None] + radau_nodes(5)[None]).ravel() / N * t[-1]
        assert out.shape == (3, N * 6, 9)
        assert np.abs(out - interp1d(t, sig, axis=1)(ti)).max() <= 1e-13
        t_host, host = interp_ws_for_collocation(t, {"s": sig[1]}, N, 5)  # the single-instance host helper behind Vehicle.interp_ws_for_collocation
        assert np.abs(t_host - ti).max() <= 1e-14 and np.abs(out[1] - host["s"]).max() <= 1e-13
