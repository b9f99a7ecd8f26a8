"""The planner classes keep the reference's construction / solve surface and run the whole
pipeline: state_ws -> dual_ws -> interp_ws_for_collocation -> setup_single_final_problem -> solve -> get_solution ->
joint_dual_ws -> solve_final_problem_obca -> interpolate_states."""
import numpy as np
import pytest

from cases import check_solution_properties
from conflict_rez_b200.control.multi_vehicle_planner import MultiVehiclePlanner
from conflict_rez_b200.control.vehicle import Vehicle, collocation_coefficients
from conflict_rez_b200.pytypes import VehicleState

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request):
    if request.param == "emu":
        return request.getfixturevalue("emu_lib"), "cpu"
    return request.getfixturevalue("cuda_lib"), "cuda:0"


def test_product_collocation_coefficients_match_oracle():
    from oracle.collocation import collocation_coefficients as ref

    for a, b in zip(collocation_coefficients(5), ref(5)):
        assert np.allclose(a, b, atol=1e-12)


def test_single_vehicle_pipeline(backend, strategy_file):
    lib, dev = backend
    v = Vehicle(rl_file_name=strategy_file, agent="vehicle_1", color={"front": (1, 0, 0), "back": (0, 1, 0)}, device=dev)
    v._lib = lib
    zu0 = v.state_ws(N=30, dt=0.1, init_offset=VehicleState(), final_heading=3 * np.pi / 2, shrink_tube=0.5, spline_ws=True)
    assert len(zu0.x) == 30 * (v.num_sets - 1) + 1
    zu0 = v.dual_ws(zu0)
    assert zu0.l.shape == (24, len(zu0.x)) and zu0.m.shape == (24, len(zu0.x))
    zu0 = v.interp_ws_for_collocation(zu0, K=5, N_per_set=5)
    v.setup_single_final_problem(zu0=zu0, init_offset=VehicleState(), final_heading=3 * np.pi / 2, K=5, N_per_set=5, shrink_tube=0.5)
    sol = v.solve_single_final_problem()
    assert sol.stats()["return_status"] == "Solve_Succeeded" and sol.stats()["t_wall_total"] > 0
    res = v.get_solution(sol)
    assert len(res.x) == v.N * 6 and len(res.l) == v.N and len(res.l[0]) == 6 and res.l[0][0].shape == (24,)
    assert np.isclose(res.t[-1], (v.N - 1 + 1.0) * res.dt)
    # interpolation reproduces the collocation nodes and holds the final state afterwards (vehicle.py:756-762)
    it = v.interpolate_states(res.t)
    assert np.allclose(it.x, res.x, atol=1e-9) and np.allclose(it.psi, res.psi, atol=1e-9)
    late = v.interpolate_states([v.N * res.dt + 1.0])
    assert np.isclose(late.x[0], res.x[-1]) and np.isclose(late.v[0], res.v[-1])
    worst = check_solution_properties(sol.problem, sol.result.z, sol.result.dt)
    assert worst["collocation"] <= 1e-2 and worst["tube"] <= 1e-2 and worst["obstacle_clearance"] >= 0.05 - 1e-2


def test_failed_solve_raises_like_opti(backend, strategy_file):
    lib, dev = backend
    v = Vehicle(rl_file_name=strategy_file, agent="vehicle_1", color={}, device=dev)
    v._lib = lib
    v.solve_options.max_iter = 2
    zu0 = v.interp_ws_for_collocation(v.dual_ws(v.state_ws(shrink_tube=0.5)))
    v.setup_single_final_problem(zu0=zu0, final_heading=3 * np.pi / 2, shrink_tube=0.5)
    with pytest.raises(RuntimeError, match="Maximum_Iterations_Exceeded"):
        v.solve_single_final_problem()


def test_multi_vehicle_planner_two_agents(backend, strategy_file):
    lib, dev = backend
    agents = ["vehicle_1", "vehicle_2"]
    offs = {a: VehicleState() for a in agents}
    offs["vehicle_1"].x.x = 0.1
    offs["vehicle_1"].e.psi = np.pi / 20
    planner = MultiVehiclePlanner(rl_file_name=strategy_file, ws_config={a: True for a in agents}, colors={a: {} for a in agents},
                                  init_offsets=offs, final_headings={"vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi}, device=dev)
    for veh in planner.vehicles.values():
        veh._lib = lib
    planner.solve_single_problems(N=30, K=5, N_per_set=5, dt=0.1, shrink_tube=0.5, dmin=0.05)
    planner.solve_final_problem_obca(K=5, N_per_set=5, shrink_tube=0.5, dmin=0.05, interp_dt=0.025)
    assert planner.final_sol.stats()["return_status"] == "Solve_Succeeded"
    assert sorted(planner.final_results) == agents
    n = len(planner.final_results["vehicle_1"].x)
    assert n == len(planner.final_results["vehicle_2"].x) and n > 100
    worst = check_solution_properties(planner.final_sol.problem, planner.final_sol.result.z, planner.final_sol.result.dt)
    assert worst["collocation"] <= 1e-2 and worst["vehicle_clearance"] >= 0.05 - 1e-2 and worst["obstacle_clearance"] >= 0.05 - 1e-2
    assert np.abs(planner.final_sol.result.z[0, 0, 0, :3] - planner.final_sol.problem.init_pose[0]).max() <= 1e-2
