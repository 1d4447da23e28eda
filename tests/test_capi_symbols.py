"""The C-ABI library loads on a CPU-only host and exports every symbol include/echem_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ech_[a-z0-9_]+)\s*\(", txt)))


def test_core_library_exports_header_symbols():
    from echemfem_b200 import build, capi
    path = build.build_core()
    lib = ctypes.CDLL(path)
    names = _declared("echem_b200.h")
    assert "ech_residual" in names and "ech_jacobian" in names and "ech_gmres" in names
    for n in names:
        assert hasattr(lib, n), "libechem_b200.so does not export %s" % n
    # the ctypes table binds exactly the declared entry points
    assert sorted(capi.EXPORTS) == names


def test_physics_module_exports_module_abi():
    import product_cases
    from echemfem_b200 import build
    from echemfem_b200.compiler import CompiledForm
    s = product_cases.AdvectionDiffusionMigrationSolver(2, 1.0)
    cf = CompiledForm(s.Form, 2, len(s._aux), 2, 1, "DG")
    mod = ctypes.CDLL(build.build_module(cf.header, cf.key))
    for n in _declared("echem_b200_module.h"):
        assert hasattr(mod, n), "physics module does not export %s" % n
    info = (ctypes.c_int32 * 64)()
    k = mod.ech_mod_describe(info, 64)
    assert k == 8 + 2 * 4
    assert list(info[:4]) == [2, 1, 2, 1]          # d, p, nf, naux
    assert list(info[8:16]) == [1, 1, 1, 1, 1, 1, 1, 1]


def test_no_cpu_fallback():
    """Without a CUDA device the product path refuses to run instead of falling back."""
    import torch
    import pytest
    import product_cases
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    s = product_cases.AdvectionDiffusionMigrationSolver(2, 1.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        s.setup_solver()
