"""CUDA assembly vs oracle: sparsity bit-exact, residual and Jacobian entries to 1e-12 (north_star tolerance).

Entry tolerance (tests/parity.py): |a - b| <= 1e-12 * max(|a|, |b|, tau_block), tau_block the
max-abs of the (row cell, column cell, field pair) element block; one scale per field for residuals.
"""
import numpy as np
import pytest
import torch

import mms_cases
import parity
import product_cases

pytestmark = pytest.mark.gpu


def _state(disc, seed=0):
    """Nodal interpolant of the exact fields + deterministic perturbation."""
    x = disc.node_coords
    ph = 7 * x[:, 0] + 3 * x[:, 1] + (5 * x[:, 2] if x.shape[1] > 2 else 0.0)
    pert = 1e-3 * np.sin(ph)     # varies in every direction: no facet sees an exactly zero normal flux (|.| kink)
    return np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.Uex(x) - pert])


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("N,vavg", [(1, 1.0), (3, 1.0), (6, 1e5), (11, 1.0)])
def test_adm_dg_q1_assembly_matches_oracle(N, vavg, variant):
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from echemfem_b200.newton import NewtonEngine

    disc = Discretization(unit_square_mesh(N), mms_cases.adm_problem(vavg))
    s = product_cases.AdvectionDiffusionMigrationSolver(N, vavg)
    eng = NewtonEngine(s)
    eng.set_variant(variant)
    u = _state(disc)
    s.u.tensor.copy_(torch.from_numpy(u).cuda())

    # sparsity: bit-exact
    indptr, indices = disc.pattern()
    rp, ci, _ = eng.csr_host()
    assert np.array_equal(rp, indptr)
    assert np.array_equal(ci, indices)

    F = eng.residual(s.u.tensor).cpu().numpy()
    parity.check_F(eng, F, disc.residual(u), "residual", scale=disc.residual_scale(u))

    eng.jacobian(s.u.tensor)
    vals = eng.vals.cpu().numpy()
    J, Jscale = disc.jacobian(u, return_scale=True)
    parity.check_J(eng, vals, (J, Jscale), "jacobian")

    # SpMV against scipy on the same matrix
    x = np.cos(np.arange(disc.ndof) * 0.37)
    y = torch.empty(disc.ndof, dtype=torch.float64, device="cuda")
    eng.spmv(torch.from_numpy(x).cuda(), y)
    parity.check_F(eng, y.cpu().numpy(), J @ x, "spmv")
    eng.set_variant(0)


def test_newton_solution_matches_oracle():
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem

    N = 8
    disc = Discretization(unit_square_mesh(N), mms_cases.adm_problem(1.0))
    u_ref, info = solve_problem(disc)
    s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0)
    s.setup_solver()
    s.solve()
    u = s.u.numpy()
    Ns = disc.ndof_scalar
    for f in range(2):
        a, b = u[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
        assert np.linalg.norm(a - b) <= 1e-8 * np.linalg.norm(b)


def _compare(disc, s, tag):
    from echemfem_b200.newton import NewtonEngine
    eng = NewtonEngine(s)
    u = _state(disc)
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    indptr, indices = disc.pattern()
    rp, ci, _ = eng.csr_host()
    assert np.array_equal(rp, indptr), tag
    assert np.array_equal(ci, indices), tag
    for variant in (0, 1):
        eng.set_variant(variant)
        F = eng.residual(s.u.tensor).cpu().numpy()
        parity.check_F(eng, F, disc.residual(u), "%s residual v%d" % (tag, variant), scale=disc.residual_scale(u))
        eng.vals.zero_()
        eng.jacobian(s.u.tensor)
        parity.check_J(eng, eng.vals.cpu().numpy(), disc.jacobian(u, return_scale=True), "%s jacobian v%d" % (tag, variant))
        # fused evaluation of both at the same state
        eng.vals.zero_()
        Ff = torch.zeros_like(s.u.tensor)
        eng.residual_jacobian(s.u.tensor, out=Ff)
        parity.check_F(eng, Ff.cpu().numpy(), disc.residual(u), "%s fused residual v%d" % (tag, variant), scale=disc.residual_scale(u))
        parity.check_J(eng, eng.vals.cpu().numpy(), disc.jacobian(u, return_scale=True), "%s fused jacobian v%d" % (tag, variant))
    eng.set_variant(0)


@pytest.mark.parametrize("p,N", [(2, 3), (3, 3)])
def test_higher_order_dg_matches_oracle(p, N):
    """DG Q2 / Q3 (cfg2 sweep P1-P3): under-integrated rule degree p+1 kept."""
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    disc = Discretization(unit_square_mesh(N), mms_cases.adm_problem(1.0, p=p))
    s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, p=p)
    _compare(disc, s, "Q%d" % p)


@pytest.mark.parametrize("p,N", [(1, 11), (2, 4), (3, 3)])
def test_bench_config_modules_match_oracle(p, N):
    """The physics modules bench.py loads (cfg2: D = (0.5e-5, 1e-5) of examples/mms_scaling.py:90-93, floats baked
    into the generated header, so these are different binaries from the D = (0.5, 1) ones above): Q1 / Q2 / Q3,
    every kernel variant, separate and fused evaluation, entry by entry."""
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    disc = Discretization(unit_square_mesh(N), mms_cases.adm_problem(1.0, D1=0.5e-5, D2=1.0e-5, p=p))
    s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, D1=0.5e-5, D2=1.0e-5, p=p)
    _compare(disc, s, "bench cfg2 Q%d" % p)


def test_graded_mesh_matches_oracle():
    """Non-uniform tensor mesh (the Bortels meshes are graded; cell volumes / facet areas vary)."""
    from oracle.mesh import tensor_mesh
    from oracle.assemble import Discretization
    from echemfem_b200.mesh import TensorMesh
    ax = [np.array([0.0, 0.1, 0.35, 0.7, 1.0]), np.array([0.0, 0.2, 0.25, 0.6, 0.8, 1.0])]
    disc = Discretization(tensor_mesh(ax), mms_cases.adm_problem(10.0))
    s = product_cases.AdvectionDiffusionMigrationSolver(0, 10.0, mesh=TensorMesh(ax))
    _compare(disc, s, "graded")


def test_extruded_3d_matches_oracle():
    """3D hexahedra, DQ1 x DG1 (cfg4 element); side walls carry ids 1-4, top/bottom are natural."""
    from oracle.mesh import unit_cube_mesh
    from oracle.assemble import Discretization
    from echemfem_b200.mesh import UnitSquareMesh, ExtrudedMesh
    disc = Discretization(unit_cube_mesh(3, 2, 2), mms_cases.adm_problem(1.0))
    base = UnitSquareMesh(3, 2)
    mesh = ExtrudedMesh(base, 2, 0.5)
    s = product_cases.AdvectionDiffusionMigrationSolver(0, 1.0, mesh=mesh)
    _compare(disc, s, "hex")


def test_bortels_three_ion_matches_oracle():
    """BASELINE configs[0]: 3 ions (H eliminated), Butler-Volmer electrodes, inlet/outlet, graded mesh.

    Also checks the block structure: 7 of the 9 field blocks exist (no Cu-SO4 coupling)."""
    from oracle.assemble import Discretization
    from echemfem_b200.meshes import BortelsMesh
    from echemfem_b200.newton import NewtonEngine
    n = (5, 6, 4, 5)
    disc = Discretization(mms_cases.bortels_oracle_mesh(*n), mms_cases.bortels_problem())
    s = product_cases.BortelsSolver(BortelsMesh(*n))
    eng = NewtonEngine(s)
    assert eng.form.own_mask.sum() == 7 and not eng.form.own_mask[0, 1] and not eng.form.own_mask[1, 0]
    x = disc.node_coords
    pert = 1e-2 * np.sin(3 * x[:, 0] + 5 * x[:, 1])
    u = np.concatenate([1.0 + pert, 1.0 - 0.5 * pert, 0.1 + pert])
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    indptr, indices = disc.pattern()
    rp, ci, _ = eng.csr_host()
    assert np.array_equal(rp, indptr) and np.array_equal(ci, indices)
    for variant in (0, 1):
        eng.set_variant(variant)
        parity.check_F(eng, eng.residual(s.u.tensor).cpu().numpy(), disc.residual(u), "bortels residual", scale=disc.residual_scale(u))
        eng.jacobian(s.u.tensor)
        parity.check_J(eng, eng.vals.cpu().numpy(), disc.jacobian(u, return_scale=True), "bortels jacobian")
    eng.set_variant(0)


def test_bortels_newton_solution_matches_oracle():
    """setup_solver() (bulk guess + potential pre-solve) + solve() vs the oracle, 1e-8 relative per field."""
    from oracle.assemble import Discretization
    from oracle.newton import solve_problem
    from echemfem_b200.meshes import BortelsMesh
    n = (6, 8, 6, 6)
    disc = Discretization(mms_cases.bortels_oracle_mesh(*n), mms_cases.bortels_problem())
    u_ref, info = solve_problem(disc)
    s = product_cases.BortelsSolver(BortelsMesh(*n))
    s.setup_solver()
    s.solve()
    u = s.u.numpy()
    Ns = disc.ndof_scalar
    for f in range(3):
        a, b = u[f * Ns:(f + 1) * Ns], u_ref[f * Ns:(f + 1) * Ns]
        assert np.linalg.norm(a - b) <= 1e-8 * max(np.linalg.norm(b), 1e-3), "field %d" % f


def test_bortels_3d_extruded_matches_oracle():
    """BASELINE configs[3] physics (examples/bortels_threeion_extruded_3D_nondim.py): the graded plane channel
    extruded into hexes, Butler-Volmer electrodes with a current density parabolic in z, ids of the plane mesh on
    the side walls, top / bottom natural."""
    from oracle.assemble import Discretization
    from echemfem_b200.mesh import ExtrudedMesh
    from echemfem_b200.meshes import BortelsMesh
    from echemfem_b200.newton import NewtonEngine
    n, layers, hz = (3, 4, 3, 3), 2, 6.0
    disc = Discretization(mms_cases.bortels_oracle_mesh(*n, layers=layers, hz=hz), mms_cases.bortels_problem(hz=hz))
    s = product_cases.Bortels3DSolver(ExtrudedMesh(BortelsMesh(*n), layers, hz / layers), hz)
    eng = NewtonEngine(s)
    assert eng.form.own_mask.sum() == 7
    x = disc.node_coords
    pert = 1e-2 * np.sin(3 * x[:, 0] + 5 * x[:, 1] + 0.7 * x[:, 2])
    u = np.concatenate([1.0 + pert, 1.0 - 0.5 * pert, 0.1 + pert])
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    indptr, indices = disc.pattern()
    rp, ci, _ = eng.csr_host()
    assert np.array_equal(rp, indptr) and np.array_equal(ci, indices)
    for variant in (0, 1):
        eng.set_variant(variant)
        parity.check_F(eng, eng.residual(s.u.tensor).cpu().numpy(), disc.residual(u), "bortels 3D residual", scale=disc.residual_scale(u))
        eng.jacobian(s.u.tensor)
        parity.check_J(eng, eng.vals.cpu().numpy(), disc.jacobian(u, return_scale=True), "bortels 3D jacobian")
    eng.set_variant(0)
    Ff = torch.zeros_like(s.u.tensor)
    eng.vals.zero_()
    eng.residual_jacobian(s.u.tensor, out=Ff)
    parity.check_F(eng, Ff.cpu().numpy(), disc.residual(u), "bortels 3D fused residual", scale=disc.residual_scale(u))
    parity.check_J(eng, eng.vals.cpu().numpy(), disc.jacobian(u, return_scale=True), "bortels 3D fused jacobian")
