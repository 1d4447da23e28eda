"""The reference's own acceptance tests (manufactured solutions, convergence-rate inequalities) run through the
CUDA path: same solver classes (tests/product_cases.py), same loops and thresholds as

  tests/test_advection_diffusion_migration.py:64-110        (DG low / high Peclet, CG)
  tests/test_diffusion_migration_poisson.py:60-96           (Poisson-coupled, DG and CG)
  tests/test_advection_diffusion_migration_porous.py:73-96  (porous, two potentials)
  tests/test_SUPG.py:68-98                                  (CG + SUPG, low / high Peclet)

of the reference; mesh sequences are cut one or two levels short of the reference's to keep the GPU suite short
(every retained level must meet the reference's threshold)."""
import pytest

import product_cases
from echemfem_b200.fd import errornorm

pytestmark = pytest.mark.gpu


def _solve(s, **kw):
    s.solver_parameters.pop("snes_monitor", None)
    s.setup_solver(**kw)
    s.solve()
    return s


def test_convergence_low_peclet():
    errC_old = errU_old = 1e6
    for i in range(5):
        s = _solve(product_cases.AdvectionDiffusionMigrationSolver(2 ** (i + 1), 1.0))
        c1, U = s.u.subfunctions
        errC, errU = errornorm(s.C1ex, c1), errornorm(s.Uex, U)
        assert errC < 0.26 * errC_old and errU < 0.26 * errU_old
        errC_old, errU_old = errC, errU


def test_convergence_high_peclet():
    errC_old = errU_old = 1e6
    for i in range(5):
        s = _solve(product_cases.AdvectionDiffusionMigrationSolver(2 ** (i + 1), 1e5))
        c1, U = s.u.subfunctions
        errC, errU = errornorm(s.C1ex, c1), errornorm(s.Uex, U)
        assert errC < 0.36 * errC_old and errU < 0.26 * errU_old          # p + 1/2 convergence
        errC_old, errU_old = errC, errU


def test_convergence_low_peclet_CG():
    errC_old = errU_old = 1e6
    for i in range(5):
        s = _solve(product_cases.AdvectionDiffusionMigrationSolver(2 ** (i + 1), 1.0, family="CG"))
        c1, U = s.u.subfunctions
        errC, errU = errornorm(s.C1ex, c1), errornorm(s.Uex, U)
        assert errC < 0.26 * errC_old and errU < 0.26 * errU_old
        errC_old, errU_old = errC, errU


@pytest.mark.parametrize("family,thr", [("DG", 0.26), ("CG", 0.27)])
def test_convergence_poisson(family, thr):
    old = [1e6] * 3
    for i in range(4):
        s = _solve(product_cases.DiffusionMigrationPoissonSolver(2 ** (i + 1), family=family))
        c1, c2, U = s.u.subfunctions
        err = [errornorm(s.C1ex, c1), errornorm(s.C2ex, c2), errornorm(s.Uex, U)]
        assert all(e < thr * o for e, o in zip(err, old))
        old = err


def test_convergence_porous():
    err_old = 1e6
    for i in range(5):
        s = _solve(product_cases.PorousSolver(2 ** (i + 1), 1.0))
        c1, U1, U2 = s.u.subfunctions
        err = errornorm(s.C1ex, c1) + errornorm(s.U1ex, U1) + errornorm(s.U2ex, U2)
        assert err < 0.29 * err_old
        err_old = err


@pytest.mark.parametrize("D,first", [(1.0, 1), (1e-4, 2)])
def test_convergence_SUPG(D, first):
    err_old = 1e6
    for i in range(4):
        s = _solve(product_cases.AdvectionDiffusionSolver(2 ** (i + first), D), initial_guess=False)
        c1, c2 = s.u.subfunctions
        err = errornorm(s.C1ex, c1) + errornorm(s.C2ex, c2)
        assert err < 0.26 * err_old
        err_old = err


def test_convergence_finite_size_CG():
    """tests/test_diffusion_finite_size.py:54-63."""
    err_old = 1e6
    for i in range(4):
        s = _solve(product_cases.DiffusionFiniteSizeSolver(2 ** (i + 1), family="CG"))
        c1, c2 = s.u.subfunctions
        err = errornorm(s.C1ex, c1) + errornorm(s.C2ex, c2)
        assert err < 0.26 * err_old
        err_old = err


@pytest.mark.parametrize("family,thr", [("DG", 0.29), ("CG", 0.26)])
def test_convergence_cylindrical(family, thr):
    """tests/test_diffusion_cylindrical.py:44-65."""
    err_old = 1e6
    for i in range(4):
        s = _solve(product_cases.DiffusionCylindricalSolver(2 ** (i + 1), family=family))
        c1, c2 = s.u.subfunctions
        err = errornorm(s.C1ex, c1) + errornorm(s.C2ex, c2)
        assert err < thr * err_old
        err_old = err


@pytest.mark.parametrize("family,D,first,thr,extruded", [
    ("DG", 1.0, 1, 0.29, False), ("DG", 1e-3, 1, 0.29, False), ("CG", 1.0, 1, 0.29, False), ("CG", 1.0 / 20.0, 2, 0.62, False),
    ("DG", 1.0, 1, 0.29, True), ("CG", 1.0, 1, 0.29, True)])
def test_convergence_advection_diffusion(family, D, first, thr, extruded):
    """tests/test_advection_diffusion.py:68-145: passive transport, DG upwind / CG, quadrilaterals and extruded hexes
    (three levels on hexes, as in the reference)."""
    err_old = 1e6
    for i in range(3 if extruded else 4):
        s = _solve(product_cases.ScalarAdvectionDiffusionSolver(2 ** (i + first), D, extruded=extruded, family=family))
        c1, c2 = s.u.subfunctions
        err = errornorm(s.C1ex, c1) + errornorm(s.C2ex, c2)
        assert err < thr * err_old
        err_old = err
