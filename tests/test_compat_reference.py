"""The reference's own test modules, loaded UNMODIFIED through the compatibility shim (build container only:
/root/reference does not exist on the GPU box, where this test skips)."""
import importlib.util
import os

import pytest

REF = "/root/reference/tests"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("fname,cls,args", [
    ("test_advection_diffusion_migration.py", "AdvectionDiffusionMigrationSolver", (4, 1.0)),
    ("test_advection_diffusion_migration_porous.py", "AdvectionDiffusionMigrationSolver", (4, 1.0)),
    ("test_SUPG.py", "AdvectionDiffusionSolver", (4, 1e-3)),
    ("test_heat_equation.py", "HeatEquationSolver", (4,)),
])
def test_reference_test_classes_build_through_the_shim(fname, cls, args):
    import echemfem_b200.compat as compat
    compat.install(force=True)
    from echemfem_b200.compiler import CompiledForm
    spec = importlib.util.spec_from_file_location("ref_" + fname[:-3], os.path.join(REF, fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                       # defines classes and test functions only
    solver = getattr(mod, cls)(*args)
    V = solver.W.scalar
    cf = CompiledForm(solver.Form, solver.W.num_sub_spaces(), len(solver._aux), solver.mesh.dim, V.degree, V.family)
    assert "cell_jac" in cf.header and cf.own_mask.any()
    # the module's own test functions exist and would drive setup_solver()/solve() on a GPU host
    assert any(n.startswith("test_") for n in dir(mod))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_unmodified_reference_class_gives_the_same_kernels_as_ours():
    """Same generated integrands (same module cache key) from the reference's file and from tests/product_cases.py."""
    import echemfem_b200.compat as compat
    compat.install(force=True)
    from echemfem_b200.compiler import CompiledForm
    import product_cases
    spec = importlib.util.spec_from_file_location("ref_adm", os.path.join(REF, "test_advection_diffusion_migration.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a = mod.AdvectionDiffusionMigrationSolver(4, 1.0)
    b = product_cases.AdvectionDiffusionMigrationSolver(4, 1.0)
    ka = CompiledForm(a.Form, 2, len(a._aux), 2, 1, "DG").key
    kb = CompiledForm(b.Form, 2, len(b._aux), 2, 1, "DG").key
    assert ka == kb
