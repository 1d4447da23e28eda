"""Continuous (CG) spaces on the device vs the oracle: node-owner gather assembly, SUPG, strong Dirichlet BCs."""
import numpy as np
import pytest
import torch

import mms_cases
import parity
import product_cases

pytestmark = pytest.mark.gpu


def _compare_raw(disc, s, u, tag):
    from echemfem_b200.newton import NewtonEngine
    eng = NewtonEngine(s)
    assert np.allclose(s.V.node_coords(), disc.node_coords)            # same node numbering as the oracle
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    indptr, indices = disc.pattern()
    rp, ci, _ = eng.csr_host()
    assert np.array_equal(rp, indptr), tag
    assert np.array_equal(ci, indices), tag
    parity.check_F(eng, eng.residual(s.u.tensor).cpu().numpy(), disc.residual(u), tag + " residual", scale=disc.residual_scale(u))
    eng.jacobian(s.u.tensor)
    parity.check_J(eng, eng.vals.cpu().numpy(), disc.jacobian(u, return_scale=True), tag + " jacobian")
    eng.vals.zero_()
    Ff = torch.zeros_like(s.u.tensor)
    eng.residual_jacobian(s.u.tensor, out=Ff)
    parity.check_F(eng, Ff.cpu().numpy(), disc.residual(u), tag + " fused residual", scale=disc.residual_scale(u))
    parity.check_J(eng, eng.vals.cpu().numpy(), disc.jacobian(u, return_scale=True), tag + " fused jacobian")
    return eng


@pytest.mark.parametrize("p", [1, 2])
def test_cg_electroneutral_assembly_matches_oracle(p):
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    N = 4
    disc = Discretization(unit_square_mesh(N), mms_cases.adm_problem(2.0, family="CG", p=p))
    s = product_cases.AdvectionDiffusionMigrationSolver(N, 2.0, family="CG", p=p)
    x = disc.node_coords
    pert = 1e-2 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u = np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.Uex(x) - pert])
    _compare_raw(disc, s, u, "CG Q%d" % p)


@pytest.mark.parametrize("D", [1.0, 1e-3])
def test_cg_supg_assembly_matches_oracle(D):
    """SUPG terms (solver.py:1655-1668): non-isotropic g3 = delta flow (x) flow, Pe switch."""
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    N = 5
    disc = Discretization(unit_square_mesh(N), mms_cases.supg_problem(D))
    s = product_cases.AdvectionDiffusionSolver(N, D)
    x = disc.node_coords
    pert = 1e-2 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u = np.concatenate([mms_cases.supg_C1ex(x) + pert, mms_cases.supg_C2ex(x) - pert])
    _compare_raw(disc, s, u, "SUPG D=%g" % D)


def test_cg_newton_solutions_match_oracle():
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    N = 8
    for make_o, make_p in (
            (lambda: mms_cases.adm_problem(1.0, family="CG"),
             lambda: product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, family="CG")),
            (lambda: mms_cases.supg_problem(1e-2), lambda: product_cases.AdvectionDiffusionSolver(N, 1e-2))):
        disc = Discretization(unit_square_mesh(N), make_o())
        u_ref, info = solve_problem(disc)
        s = make_p()
        s.setup_solver()
        s.solve()
        u = s.u.numpy()
        Ns = disc.ndof_scalar
        for f in range(2):
            a, b = u[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
            assert np.linalg.norm(a - b) <= 1e-8 * np.linalg.norm(b)


@pytest.mark.parametrize("family", ["DG", "CG"])
def test_poisson_coupled_matches_oracle(family):
    """Poisson-coupled Nernst-Planck (potential_poisson_form, solver.py:2048-2234), three fields."""
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    N = 4
    disc = Discretization(unit_square_mesh(N), mms_cases.poisson_problem(family))
    s = product_cases.DiffusionMigrationPoissonSolver(N, family=family)
    x = disc.node_coords
    pert = 1e-2 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u = np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.poisson_C2ex(x) - pert, mms_cases.Uex(x) + 0.5 * pert])
    _compare_raw(disc, s, u, "poisson " + family)
    u_ref, info = solve_problem(disc)
    s2 = product_cases.DiffusionMigrationPoissonSolver(N, family=family)
    s2.setup_solver()
    s2.solve()
    un = s2.u.numpy()
    Ns = disc.ndof_scalar
    for f in range(3):
        a, b = un[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
        assert np.linalg.norm(a - b) <= 1e-8 * np.linalg.norm(b)


@pytest.mark.parametrize("family", ["DG", "CG"])
@pytest.mark.parametrize("variant", ["robin", "applied liquid current", "liquid conductivity"])
def test_poisson_boundary_branches_match_oracle(family, variant):
    """The Poisson branches no reference test reaches (echemfem/solver.py:2085-2093 "liquid conductivity",
    :2219-2232 applied current density / Robin "gap capacitance"): residual, Jacobian and Newton solution."""
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    N = 4
    disc = Discretization(unit_square_mesh(N), mms_cases.poisson_problem(family, variant))
    s = product_cases.DiffusionMigrationPoissonSolver(N, family=family, variant=variant)
    x = disc.node_coords
    pert = 1e-2 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u = np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.poisson_C2ex(x) - pert, mms_cases.Uex(x) + 0.5 * pert])
    _compare_raw(disc, s, u, "poisson %s %s" % (variant, family))
    u_ref, info = solve_problem(disc)
    s2 = product_cases.DiffusionMigrationPoissonSolver(N, family=family, variant=variant)
    s2.setup_solver()
    s2.solve()
    un = s2.u.numpy()
    Ns = disc.ndof_scalar
    for f in range(3):
        a, b = un[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
        assert np.linalg.norm(a - b) <= 1e-8 * np.linalg.norm(b), (variant, family, f)


def test_porous_applied_solid_current_matches_oracle():
    """"applied solid current" on one side of the solid-potential equation (solver.py:2219-2221)."""
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    N = 4
    disc = Discretization(unit_square_mesh(N), mms_cases.porous_problem(2.0, solid_current=True))
    s = product_cases.PorousSolver(N, 2.0, solid_current=True)
    x = disc.node_coords
    pert = 1e-2 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u = np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.Uex(x) - pert, mms_cases.Uex(x) + 0.5 * pert])
    _compare_raw(disc, s, u, "porous, applied solid current")


def test_porous_two_potentials_matches_oracle():
    """Porous electrode model (cfg5 ingredients): liquid + solid potential, coupled pre-solve."""
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    N = 4
    disc = Discretization(unit_square_mesh(N), mms_cases.porous_problem(2.0))
    s = product_cases.PorousSolver(N, 2.0)
    x = disc.node_coords
    pert = 1e-2 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u = np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.Uex(x) - pert, mms_cases.Uex(x) + 0.5 * pert])
    _compare_raw(disc, s, u, "porous")
    u_ref, info = solve_problem(disc)
    s2 = product_cases.PorousSolver(N, 2.0)
    s2.setup_solver()
    s2.solve()
    un = s2.u.numpy()
    Ns = disc.ndof_scalar
    for f in range(3):
        a, b = un[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
        assert np.linalg.norm(a - b) <= 1e-8 * np.linalg.norm(b)


@pytest.mark.parametrize("kind", ["BMCSL", "finite size"])
def test_finite_size_mass_action_matches_oracle(kind):
    """cfg5 ingredients on the device: finite-size chemical potential (solver.py:1543-1636), homogeneous
    reactions from homog_params (solver.py:1222-1274), Poisson coupling; CG Q1, four fields."""
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    N = 4
    disc = Discretization(unit_square_mesh(N), mms_cases.finite_size_problem(kind))
    s = product_cases.FiniteSizeSolver(N, kind)
    x = disc.node_coords
    pert = 1e-2 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u = np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3 + pert, np.cos(x[:, 0]) - np.sin(x[:, 1]) + 3 - pert,
                        np.sin(x[:, 0]) * np.cos(x[:, 1]) + 2 + 0.5 * pert, mms_cases.Uex(x) + 0.5 * pert])
    _compare_raw(disc, s, u, "finite size " + kind)
    if kind == "BMCSL":
        u_ref, info = solve_problem(disc)
        s2 = product_cases.FiniteSizeSolver(N, kind)
        s2.setup_solver()
        s2.solve()
        un = s2.u.numpy()
        Ns = disc.ndof_scalar
        for f in range(4):
            a, b = un[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
            assert np.linalg.norm(a - b) <= 1e-8 * np.linalg.norm(b)


@pytest.mark.parametrize("SUPG", [False, True])
def test_gupta_cg_five_species_matches_oracle(SUPG):
    """BASELINE configs[2] physics (examples/gupta_advection_migration.py) on a small boundary-layer mesh: CG,
    K eliminated, two mass-action reactions, constant-current electrode fluxes, shear flow; with and without
    SUPG (the north star's variant)."""
    from oracle.assemble import Discretization
    from echemfem_b200.newton import NewtonEngine
    disc = Discretization(mms_cases.gupta_oracle_mesh(), mms_cases.gupta_problem(SUPG=SUPG))
    s = product_cases.GuptaSolver(SUPG=SUPG)
    x = disc.node_coords
    ph = np.sin(7e2 * x[:, 0] + 3e3 * x[:, 1])
    bulk = [b for _, _, b, _ in product_cases.GUPTA_SPECIES[:4]]
    u = np.concatenate([b * (1.0 + 0.05 * ph * (1 + 0.1 * k)) for k, b in enumerate(bulk)] + [1e-3 * ph])
    _compare_raw(disc, s, u, "gupta SUPG=%s" % SUPG)     # residuals are compared field by field (tests/parity.py)


@pytest.mark.parametrize("p", [1, 2])
def test_catalyst_layer_full_electroneutrality_matches_oracle(p):
    """BASELINE configs[4] physics (examples/catalyst_layer_Cu_full.py): porous electrode, six species + explicit
    electroneutrality constraint with shifted test functions, liquid + solid potential, volumetric Tafel kinetics,
    gas-side Dirichlet value, Robin exchange with the bulk; CG Q1 and Q2 (the example uses p = 2)."""
    from oracle.mesh import tensor_mesh
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    N = 4
    mesh = tensor_mesh([np.linspace(0.0, 1.0, N + 1), np.linspace(0.0, 0.5, N // 2 + 1)])
    disc = Discretization(mesh, mms_cases.catalyst_layer_problem(p))
    s = product_cases.CatalystLayerSolver(N, p=p)
    x = disc.node_coords
    ph = np.sin(5 * x[:, 0] + 3 * x[:, 1])
    bulk = [b for _, _, _, b, _, _ in product_cases.CL_SPECIES]
    u = np.concatenate([b * (1.0 + 0.05 * ph * (1 + 0.1 * k)) for k, b in enumerate(bulk)] +
                       [-0.02 * (1 + ph), -0.3 + 0.02 * ph])
    _compare_raw(disc, s, u, "catalyst layer p=%d" % p)
    if p == 1:
        u_ref, info = solve_problem(disc)
        s2 = product_cases.CatalystLayerSolver(N, p=p)
        s2.setup_solver(initial_solve=False)
        s2.solve()
        un = s2.u.numpy()
        Ns = disc.ndof_scalar
        for f in range(8):
            a, b = un[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
            assert np.linalg.norm(a - b) <= 1e-8 * max(np.linalg.norm(b), 1e-2), "field %d" % f


def test_bmcsl_1d_interval_matches_oracle():
    """examples/BMCSL.py physics in 1D (d = 1 kernels): CG P2 on an interval with a boundary layer at the charged
    wall; assembly entries and the converged Newton solution."""
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    disc = Discretization(mms_cases.bmcsl_1d_oracle_mesh(), mms_cases.bmcsl_1d_problem())
    s = product_cases.BMCSL1DSolver()
    x = disc.node_coords[:, 0]
    u = np.concatenate([1.0 + 0.05 * np.sin(3 * x), 0.9 - 0.04 * np.sin(2 * x), 0.1 + 0.01 * np.cos(4 * x), 0.2 * x * x + 0.1 * x])        # U'(0) != 0: off the kink of conditional(flow.n)
    _compare_raw(disc, s, u, "BMCSL 1D")
    u_ref, info = solve_problem(disc)
    s2 = product_cases.BMCSL1DSolver()
    s2.setup_solver()
    s2.solve()
    un = s2.u.numpy()
    Ns = disc.ndof_scalar
    for f in range(4):
        a, b = un[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
        assert np.linalg.norm(a - b) <= 1e-8 * max(np.linalg.norm(b), 1e-2), "field %d" % f
