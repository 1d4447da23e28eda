"""TransientEchemSolver (backward Euler, solver.py:2629-2713) on the device vs the oracle, and the stats CSV."""
import os

import numpy as np
import pytest
import torch

import mms_cases
import product_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("family", ["DG", "CG"])
def test_backward_euler_step_matches_oracle(family):
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    from echemfem_b200.newton import NewtonEngine
    N, t1, dt = 4, 0.3, 0.1
    s = product_cases.HeatEquationSolver(N, family=family)
    s.setup_solver()
    x = s.V.node_coords()
    u_old = np.concatenate([np.sin(x[:, 0]) + np.cos(x[:, 1]), np.cos(x[:, 0]) + np.sin(x[:, 1])]) * 1.1
    s.u_old.tensor.copy_(torch.from_numpy(u_old).cuda())
    prob, c1, c2 = mms_cases.heat_problem(t1, dt, u_old, family)
    disc = Discretization(unit_square_mesh(N), prob)
    # one time step through the public API
    s.solve([t1 - dt, t1])
    u_ref, info = solve_problem(disc, u0=np.concatenate([c1(disc.node_coords), c2(disc.node_coords)]) * 0 + u_old)
    u = s.u.numpy()
    assert np.linalg.norm(u - u_ref) <= 1e-8 * np.linalg.norm(u_ref)
    # u_old.assign(u) happened (solver.py:2709)
    assert np.array_equal(s.u_old.numpy(), u)


def test_heat_equation_converges():
    """Reference acceptance test tests/test_heat_equation.py:45-56 (first three refinements)."""
    from echemfem_b200.fd import errornorm
    err_old = 1e6
    for i in range(3):
        solver = product_cases.HeatEquationSolver(2 ** (i + 1))
        solver.setup_solver()
        times = np.linspace(0, 1, 1 + 2 ** (2 * (i + 1)))
        solver.solve(times)
        c1, c2 = solver.u.subfunctions
        err = errornorm(solver.C1ex, c1) + errornorm(solver.C2ex, c2)
        assert err < 0.27 * err_old
        err_old = err


def test_stats_csv(tmp_path):
    """Same CSV columns as the reference (solver.py:761-779)."""
    import pandas
    path = str(tmp_path / "stats" / "data.csv")
    s = product_cases.AdvectionDiffusionMigrationSolver(4, 1.0)
    s.stats_file = path
    s.csolver = "asm"
    s.setup_solver()
    s.solve()
    df = pandas.read_csv(path)
    assert list(df.columns) == ["num_processes", "num_cells", "dimension", "degreeC", "degreeU", "dofs", "name",
                                "snes_its", "ksp_its", "SNESSolve", "KSPSolve", "PCSetUp", "PCApply",
                                "JacobianEval", "FunctionEval", "mesh_name", "csolver"]
    assert df["dofs"][0] == 128 and df["snes_its"][0] >= 1 and df["csolver"][0] == "asm"
