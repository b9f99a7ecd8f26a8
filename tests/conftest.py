import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def strategy_file():
    from conflict_rez_b200.control.strategy import write_strategy

    fn = os.path.join(tempfile.mkdtemp(), "4v_rl_traj")
    write_strategy(fn)
    return fn


@pytest.fixture(scope="session")
def emu_lib():
    """Developer host emulation (single-thread CPU build of the kernels) -- used by CPU-side tests of the host logic
    and of the algorithm; never used by the product path."""
    from conflict_rez_b200 import solver

    path = os.path.join(ROOT, "tools", "host_emu", "libobca_hostemu.so")
    src = os.path.join(ROOT, "conflict_rez_b200", "csrc")
    stale = not os.path.exists(path) or any(os.path.getmtime(os.path.join(src, f)) > os.path.getmtime(path) for f in os.listdir(src))
    if stale:
        try:
            subprocess.run(["sh", os.path.join(ROOT, "tools", "host_emu", "build.sh")], check=True, capture_output=True)
        except Exception as e:  # pragma: no cover
            pytest.skip("host emulation could not be built: %s" % e)
    return solver.load_library(path)


@pytest.fixture(scope="session")
def cuda_lib():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from conflict_rez_b200 import solver

    return solver.load_library()


def make_case(strategy_file, agents, **kw):
    from conflict_rez_b200.control.scenario import build_guess, build_problem

    prob = build_problem(strategy_file, list(agents), **kw)
    return prob, build_guess(prob, strategy_file, list(agents))
