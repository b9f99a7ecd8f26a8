"""Two library handles on two GPUs inside ONE process (the device guard of every C-ABI entry, csrc/obca_api.cu ``DeviceGuard``).

The bench and the other GPU tests run one process per GPU with ``torch.cuda.set_device`` first, so a missing ``cudaSetDevice`` in
an entry point would go unnoticed there.  Here torch's current device stays 0 throughout while a second handle lives on ``cuda:1``;
the calls on the two handles are interleaved and both must reproduce the single-device result bit for bit.  Needs two GPUs
(``gpurun --gpus 2``); skipped on a single-GPU box.
"""
import numpy as np
import pytest

from cases import load_golden

from conflict_rez_b200.solver import ObcaSolver, SolveOptions


@pytest.mark.gpu
def test_two_handles_on_two_devices_interleaved(cuda_lib):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    prob, guess, _ = load_golden("single_vehicle_1")
    prob2, guess2, _ = load_golden("joint_vehicle_1_2")
    opts = SolveOptions(tol=1e-6, constr_viol_tol=1e-6)
    torch.cuda.set_device(0)
    ref = ObcaSolver(prob, opts, device="cuda:0", lib=cuda_lib).solve(guess)
    ref2 = ObcaSolver(prob2, opts, device="cuda:0", lib=cuda_lib).solve(guess2)
    # interleaved life cycles: create on 1, create on 0, upload / run / fetch alternating, torch's current device untouched
    s1 = ObcaSolver(prob, opts, device="cuda:1", lib=cuda_lib)
    s0 = ObcaSolver(prob2, opts, device="cuda:0", lib=cuda_lib)
    assert torch.cuda.current_device() == 0
    d1 = s1.upload(guess)
    d0 = s0.upload(guess2)
    s1.set_inputs(d1)
    s0.set_inputs(d0)
    s1.run()
    s0.run()
    assert torch.cuda.current_device() == 0
    r0 = s0.solve(guess2)
    r1 = s1.solve(guess)
    assert torch.cuda.current_device() == 0
    for got, want in ((r1, ref), (r0, ref2)):
        assert got.status[0] == want.status[0] == 0
        assert got.iters[0] == want.iters[0]
        np.testing.assert_array_equal(got.z, want.z)
        np.testing.assert_array_equal(got.dt, want.dt)
        np.testing.assert_array_equal(got.lam, want.lam)
    s1.close()
    s0.close()
    # a handle created while ANOTHER device is current must still live on its own device
    torch.cuda.set_device(1)
    try:
        r = ObcaSolver(prob, opts, device="cuda:0", lib=cuda_lib).solve(guess)
        assert torch.cuda.current_device() == 1
    finally:
        torch.cuda.set_device(0)
    np.testing.assert_array_equal(r.z, ref.z)
