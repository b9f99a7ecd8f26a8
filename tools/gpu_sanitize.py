"""Developer script: one short pass through every kernel of the library for compute-sanitizer.

    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py single joint4 mpc ws traj
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py joint2
    compute-sanitizer --tool synccheck python tools/gpu_sanitize.py joint2 mpc
    compute-sanitizer --tool initcheck python tools/gpu_sanitize.py single

The IPM is cut off after a few iterations (``max_iter``): every phase of an iteration (evaluation, local block elimination,
null space, Riccati, back-substitution, flat passes, line search) runs, the solve is not meant to converge.  The cases are the
golden problems of tests/golden (no oracle involved) and a few control steps of the closed loop.
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from cases import load_golden
from conflict_rez_b200.solver import ObcaSolver, SolveOptions

IT = int(os.environ.get("SANITIZE_ITERS", "6"))
cases = sys.argv[1:] or ["single", "joint2", "joint4", "mpc", "ws", "traj"]


def golden_case(name, iters=IT, refine=0):
    prob, guess, _ = load_golden(name)
    opts = SolveOptions(tol=1e-8, constr_viol_tol=1e-8, max_iter=iters)
    if refine:
        opts.refine_steps = refine
    sv = ObcaSolver(prob, opts, device="cuda:0")
    t0 = time.perf_counter()
    res = sv.solve(guess)
    print("%-28s status %-28s iters %3d  cviol %.2e  %.1f s" % (name, res.return_status(0), int(res.iters[0]), float(res.cviol[0]), time.perf_counter() - t0), flush=True)
    sv.close()


for c in cases:
    if c == "single":
        golden_case("single_vehicle_1")
        golden_case("single_vehicle_2_free_heading")
    elif c == "joint2":
        golden_case("joint_vehicle_1_2")
    elif c == "joint4":
        golden_case("joint_vehicle_0_1_2_3")
    elif c == "mpc":
        from conflict_rez_b200.control.strategy import write_strategy
        from conflict_rez_b200.control.vehicle_follower import DeviceMpcLoop, MultiDistributedFollower
        from conflict_rez_b200.pytypes import VehicleState

        AGENTS = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
        HEADINGS = {"vehicle_0": 0.0, "vehicle_1": 3 * np.pi / 2, "vehicle_2": np.pi, "vehicle_3": np.pi / 2}
        fn = os.path.join(tempfile.mkdtemp(), "4v")
        write_strategy(fn)
        np.random.seed(0)
        m = MultiDistributedFollower(fn, {a: True for a in AGENTS}, {a: {} for a in AGENTS}, {a: VehicleState() for a in AGENTS}, HEADINGS, device="cuda:0")
        m.setup_multi_vehicles()
        loop = DeviceMpcLoop(m)
        loop.run(int(os.environ.get("SANITIZE_MPC_STEPS", "3")))
        ex = loop.export()
        print("mpc closed loop: statuses", ex["status"].tolist(), "iters", ex["iters"].tolist(), flush=True)
    elif c == "ws":
        from conflict_rez_b200.control.batch_planner import prepare_joint_batch, random_init_offsets
        from conflict_rez_b200.control.strategy import write_strategy

        fn = os.path.join(tempfile.mkdtemp(), "4v")
        write_strategy(fn)
        agents = ["vehicle_0", "vehicle_1", "vehicle_2", "vehicle_3"]
        # the whole warm-start pipeline (state_ws Euler NLP, interp_ws, dual_ws, joint_dual_ws, four single-vehicle solves), two instances
        plan = prepare_joint_batch(fn, agents, random_init_offsets(2, 4, seed=0), SolveOptions(tol=1e-2, constr_viol_tol=1e-2, max_iter=int(os.environ.get("SANITIZE_WS_ITERS", "12"))))
        print("warm-start pipeline:", plan.timing, flush=True)
    elif c == "traj":
        import subprocess

        # the trajectory kernels (interpolation, reference lookup, horizon shift, plant step) through their own GPU tests
        import pytest

        rc = pytest.main(["-q", "-x", "-m", "gpu", os.path.join(ROOT, "tests", "test_trajectory_kernels.py"), os.path.join(ROOT, "tests", "test_warmstart.py")])
        print("trajectory / warm-start kernel tests rc", int(rc), flush=True)
    else:
        raise SystemExit("unknown case " + c)
print("done", flush=True)
