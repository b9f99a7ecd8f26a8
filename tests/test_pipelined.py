"""PipelinedSolver (two handles on two streams, consecutive batches overlapping) returns exactly what ObcaSolver.solve returns."""
import numpy as np
import pytest

from cases import load_golden

from conflict_rez_b200.solver import ObcaSolver, PipelinedSolver, SolveOptions


@pytest.mark.gpu
def test_pipelined_batches_equal_plain_solves(cuda_lib):
    import torch

    prob, guess, _ = load_golden("joint_vehicle_1_2")
    opts = SolveOptions(tol=1e-6, constr_viol_tol=1e-6)
    ref = ObcaSolver(prob, opts, device="cuda:0", lib=cuda_lib).solve(guess)
    assert ref.status[0] == 0
    # a second, different batch: the same problem from a perturbed start
    g2 = type(guess)(guess.z + 1e-3, guess.lam, guess.mu, guess.dt, guess.pair_lam, guess.pair_mu, guess.pair_s)
    ref2 = ObcaSolver(prob, opts, device="cuda:0", lib=cuda_lib).solve(g2)
    pipe = PipelinedSolver(prob, opts, device="cuda:0", lib=cuda_lib, depth=2)
    out = pipe.solve_many([guess, g2, guess, g2, guess])
    for k, r in enumerate(out):
        want = ref if k % 2 == 0 else ref2
        assert r.status[0] == want.status[0] and r.iters[0] == want.iters[0]
        np.testing.assert_array_equal(r.z, want.z)
        np.testing.assert_array_equal(r.dt, want.dt)
        np.testing.assert_array_equal(r.pair_lam, want.pair_lam)
    # device-resident form: every step (alternating handles) reports the same statistics
    d = pipe.solvers[0].upload(guess)
    torch.cuda.synchronize()
    stats = pipe.run_resident(d, 4)
    torch.cuda.synchronize()
    for st, it, dbl in stats:
        assert int(st[0]) == int(ref.status[0]) and int(it[0]) == int(ref.iters[0])
        assert float(dbl[0][0]) == float(ref.obj[0])
    assert pipe.launch_count >= 9
    pipe.close()


def test_pipelined_solver_refuses_cpu_device():
    prob, _, _ = load_golden("single_vehicle_1")
    with pytest.raises(RuntimeError, match="CUDA device"):
        PipelinedSolver(prob, SolveOptions(), device="cpu")
