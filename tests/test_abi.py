"""The C-ABI library exports every symbol ``include/obca.h`` declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "obca.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(obca_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from conflict_rez_b200 import solver

    assert declared_symbols() == sorted(solver.EXPORTS)


def test_cuda_library_exports_every_declared_symbol():
    from conflict_rez_b200 import solver

    path = solver.default_library_path()
    if not os.path.exists(path):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.obca_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.obca_version()


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly (ObcaSolver and obca_create)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from conflict_rez_b200 import solver

    lib = solver.load_library()
    dims = solver.ObcaDims(batch=1, V=1, O=1, K=5, n_per_set=5)
    dims.n_sets[0] = 3
    h = ctypes.c_void_p()
    assert lib.obca_create(ctypes.byref(dims), None, 0, ctypes.byref(h)) < 0
    assert b"no CUDA device" in lib.obca_last_error()
    from cases import load_golden

    prob, guess, _ = load_golden("single_vehicle_1")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        solver.ObcaSolver(prob)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        solver.ObcaSolver(prob, device="cpu")


def test_missing_library_is_an_error(tmp_path):
    from conflict_rez_b200 import solver

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        solver.load_library(str(tmp_path / "libobca_b200.so"))
