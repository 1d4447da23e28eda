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


EXAMPLES = "/root/reference/examples"


@pytest.mark.skipif(not os.path.isdir(EXAMPLES), reason="reference tree not present")
@pytest.mark.parametrize("script,nf,dim,family", [
    ("bortels_threeion_nondim.py", 3, 2, "DG"),          # BASELINE configs[0]: Mesh('...msh') built from the .geo
    ("gupta_advection_migration.py", 5, 2, "CG"),        # configs[2]: homog_params, RectangleBoundaryLayerMesh
    ("catalyst_layer_Cu_full.py", 8, 1, "CG"),           # configs[4]: porous, electroneutrality full, moved vertices
    ("BMCSL.py", 4, 1, "CG"),                            # finite-size Poisson-Nernst-Planck, IntervalBoundaryLayerMesh
])
def test_reference_example_scripts_run_unmodified_up_to_the_device(script, nf, dim, family):
    """The reference's example scripts, executed UNMODIFIED: imports (`firedrake`, `echemfem`, mpi4py / petsc4py /
    matplotlib stand-ins), mesh construction, the solver class and its weak form all go through this package;
    `setup_solver()` is intercepted after the form has been compiled to device code (no GPU in this container)."""
    import contextlib
    import io
    import echemfem_b200.compat as compat
    compat.install(force=True)
    from echemfem_b200.compiler import CompiledForm
    from echemfem_b200 import solver as S
    seen = {}

    def compile_only(self, *a, **k):
        V = self.W.scalar
        cf = CompiledForm(self.Form, self.W.num_sub_spaces(), len(self._aux), self.mesh.dim, V.degree, V.family)
        seen.update(nf=self.W.num_sub_spaces(), dim=self.mesh.dim, family=V.family, header=cf.header)
        raise SystemExit("compiled")
    orig = S.EchemSolver.setup_solver
    S.EchemSolver.setup_solver = compile_only
    cwd = os.getcwd()
    path = os.path.join(EXAMPLES, script)
    try:
        os.chdir(EXAMPLES)
        with open(path) as f:
            src = f.read()
        with contextlib.redirect_stdout(io.StringIO()):
            with pytest.raises(SystemExit, match="compiled"):
                exec(compile(src, path, "exec"), {"__name__": "__main__", "__file__": path})
    finally:
        os.chdir(cwd)
        S.EchemSolver.setup_solver = orig
    assert (seen["nf"], seen["dim"], seen["family"]) == (nf, dim, family)
    assert "cell_jac" in seen["header"]
