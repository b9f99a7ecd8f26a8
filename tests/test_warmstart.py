"""Device dual warm starts (obca_dual_ws / obca_joint_dual_ws) against the host closed form of
control/warmstart.py, which itself is checked against an independent geometric distance in test_data_model.py.

The reference computes these duals with one IPOPT call each (vehicle.py:233-296, multi_vehicle_planner.py:208-341);
for 4-face polytopes the maximiser is the closest-point direction, so the comparison is exact up to rounding.
"""
import numpy as np
import pytest
import torch

from conflict_rez_b200.control import warmstart
from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
from conflict_rez_b200.control.scenario import build_guess, build_problem
from conflict_rez_b200.solver import ObcaSolver, SolveOptions

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
AGENTS = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]


@pytest.fixture(params=BACKENDS)
def backend(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_lib"), "cpu"
    return request.getfixturevalue("cuda_lib"), "cuda:0"


def test_device_duals_match_host_closed_form(backend, strategy_file):
    lib, dev = backend
    prob = build_problem(strategy_file, AGENTS, init_offsets=random_init_offsets(3, 4, seed=1))
    g = build_guess(prob, strategy_file, AGENTS)  # host closed form
    sv = ObcaSolver(prob, SolveOptions(), device=dev, lib=lib)
    z = torch.from_numpy(g.z.copy()).to(dev)
    lam, mu = sv.dual_ws(z)
    pl, pm, ps = sv.joint_dual_ws(z)
    for got, want in ((lam, g.lam), (mu, g.mu), (pl, g.pair_lam), (pm, g.pair_mu), (ps, g.pair_s)):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=1e-12)
    # dual feasibility of the obstacle duals: |A' lam| = 1 and G' mu + R' A' lam = 0 wherever the node exists
    M1 = int(prob.nodes[1])
    l1, m1 = lam[:, 1, :M1].cpu().numpy(), mu[:, 1, :M1].cpu().numpy()
    Atl = np.einsum("orc,bnor->bnoc", prob.obs_A, l1)
    np.testing.assert_allclose(np.linalg.norm(Atl, axis=-1), 1.0, atol=1e-9)
    psi = g.z[:, 1, :M1, 2]
    c, s = np.cos(psi)[..., None], np.sin(psi)[..., None]
    RtAtl = np.stack([c * Atl[..., 0] + s * Atl[..., 1], -s * Atl[..., 0] + c * Atl[..., 1]], -1)
    np.testing.assert_allclose(np.einsum("rc,bnor->bnoc", prob.body_G, m1) + RtAtl, 0.0, atol=1e-9)
    # overlapping bodies (the joint warm start of vehicles that still collide): least-penetration axis, same as the host
    z2 = z.clone()
    z2[:, 1, :100, :3] = z2[:, 0, :100, :3] + torch.tensor([0.5, 0.3, 0.4], dtype=torch.float64, device=z.device)
    pl2, pm2, ps2 = sv.joint_dual_ws(z2)
    za, zb = z2[:, 0, :100].cpu().numpy(), z2[:, 1, :100].cpu().numpy()
    l_, m_, s_ = warmstart.joint_dual_ws_rect(za[..., 0], za[..., 1], za[..., 2], zb[..., 0], zb[..., 1], zb[..., 2], prob.body_G, prob.body_g)
    q = prob.pairs.index((0, 1))
    np.testing.assert_allclose(pl2[:, q, :100].cpu().numpy(), l_, atol=1e-12)
    np.testing.assert_allclose(pm2[:, q, :100].cpu().numpy(), m_, atol=1e-12)
    np.testing.assert_allclose(ps2[:, q, :100].cpu().numpy(), s_, atol=1e-12)
    # padding nodes of the shorter vehicles are zero
    assert float(lam[:, 1, M1:].abs().max()) == 0.0 and float(pl[:, q, M1:].abs().max()) == 0.0
    sv.close()


def test_device_warm_start_pipeline(backend, strategy_file):
    """prepare_joint_batch (two agents to keep the emulation fast): joint guess = single solutions + device pair duals."""
    lib, dev = backend
    agents = ["vehicle_1", "vehicle_2"]
    plan = prepare_joint_batch(strategy_file, agents, random_init_offsets(2, 4, seed=3)[:, 1:3], SolveOptions(tol=1e-2, constr_viol_tol=1e-2, max_iter=300), device=dev, lib=lib)
    plan.solver.close()
    g, prob = plan.guess, plan.problem
    for ia, r in enumerate(plan.singles):
        assert (r.status >= 0).all()
        M = int(prob.nodes[ia])
        np.testing.assert_array_equal(g.z[:, ia, :M], r.z[:, 0, :M])
    np.testing.assert_allclose(g.dt, np.mean([r.dt for r in plan.singles], axis=0), rtol=1e-15)
    m = int(prob.nodes.min())
    za, zb = g.z[:, 0, :m], g.z[:, 1, :m]
    l_, m_, s_ = warmstart.joint_dual_ws_rect(za[..., 0], za[..., 1], za[..., 2], zb[..., 0], zb[..., 1], zb[..., 2], prob.body_G, prob.body_g)
    np.testing.assert_allclose(g.pair_lam[:, 0, :m], l_, atol=1e-12)
    np.testing.assert_allclose(g.pair_s[:, 0, :m], s_, atol=1e-12)
    assert set(plan.timing) >= {"host_pose_guess_s", "device_warm_start_s"}


def test_batched_pipeline_with_the_reference_state_ws(emu_lib, strategy_file):
    """prepare_joint_batch(state_ws="euler"): Vehicle.state_ws (Euler NLP) -> interp_ws_for_collocation -> dual_ws -> single OBCA solve per
    agent, all batched on the device -- the reference chain (multi_vehicle_planner.py:68-109).  It must reach the same joint plans
    as the default pipeline (tube following in collocation form)."""
    from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
    from conflict_rez_b200.solver import SolveOptions

    agents = ["vehicle_1", "vehicle_2"]
    offs = random_init_offsets(2, 4, seed=0)[:, [1, 2]]
    out = {}
    for mode in ("collocation", "euler"):
        plan = prepare_joint_batch(strategy_file, agents, offs, SolveOptions(max_iter=600), device="cpu", lib=emu_lib, state_ws=mode)
        assert plan.timing["state_ws_failures"] == 0 and all((r.status >= 0).all() for r in plan.singles)
        out[mode] = plan.solver.solve(plan.guess)
        plan.solver.close()
        assert (out[mode].status == 0).all()
    assert np.abs(out["euler"].obj - out["collocation"].obj).max() <= 1e-4 * np.abs(out["collocation"].obj).max()
    assert np.abs(out["euler"].z[..., :3] - out["collocation"].z[..., :3]).max() <= 5e-2  # both stop at tol = 1e-2
