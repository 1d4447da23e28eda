"""Geometric multigrid preconditioner (the reference's `gmg` option set, solver.py:1190-1201) on the device."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import mms_cases
import product_cases
from echemfem_b200.mesh import UnitSquareMesh

pytestmark = pytest.mark.gpu


def _system(N, tile=None, p=1):
    from echemfem_b200.newton import NewtonEngine
    # Q1: the transport-dominated parameters of the benchmark; Q2: the diffusion-dominated ones of the MMS tests
    D1, D2 = (0.5e-5, 1.0e-5) if p == 1 else (0.5, 1.0)
    s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, D1=D1, D2=D2, p=p,
                                                        mesh=UnitSquareMesh(N, N, tile=tile))
    eng = NewtonEngine(s, nblocks=max(1, (N * N) // 16))
    x = s.V.node_coords()
    u = np.concatenate([mms_cases.C1ex(x), mms_cases.Uex(x) + 1e-3 * np.sin(7 * x[:, 0] + 3 * x[:, 1])])
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F = eng.residual(s.u.tensor).clone()
    eng.jacobian(s.u.tensor)
    rp, ci, v = eng.csr_host()
    A = sp.csr_matrix((v, ci, rp), shape=(eng.ndof, eng.ndof))
    return eng, A, F


@pytest.mark.parametrize("tile,p", [(None, 1), (4, 1)])
def test_galerkin_levels_and_vcycle_solution(tile, p):
    from echemfem_b200.multigrid import Multigrid
    eng, A, F = _system(16, tile, p)
    M = Multigrid(eng, tile=4, cells_per_block=16, coarsest_cells=16)
    assert M.nlevels == 3
    M.setup()
    # level-1 operator = P^T A P with the host prolongation
    P = sp.block_diag([M.h.P[0]] * eng.nf).tocsr()
    Ac = (P.T @ A @ P).tocsr()
    D = M.lv[1]
    got = sp.csr_matrix((D.vals.cpu().numpy(), D.colidx.cpu().numpy(), D.rowptr.cpu().numpy()), shape=Ac.shape)
    assert got.has_sorted_indices
    assert abs(got - Ac).max() <= 1e-11 * abs(Ac).max()
    # transfers (ech_dg_restrict / ech_dg_prolong_add) against the host prolongation
    D0, D1 = M.lv[0], M.lv[1]
    rng = np.random.default_rng(3)
    rf, zc = rng.standard_normal(D0.ndof), rng.standard_normal(D1.ndof)
    z0 = rng.standard_normal(D0.ndof)
    rc_d = torch.zeros(D1.ndof, dtype=torch.float64, device="cuda")
    M._restrict(D0, D1, torch.from_numpy(rf).cuda(), rc_d)
    np.testing.assert_allclose(rc_d.cpu().numpy(), P.T @ rf, rtol=0, atol=1e-13 * np.abs(rf).max() * 16)
    z_d = torch.from_numpy(z0).cuda()
    M._prolong_add(D0, D1, torch.from_numpy(zc).cuda(), z_d)
    np.testing.assert_allclose(z_d.cpu().numpy(), z0 + P @ zc, rtol=0, atol=1e-13 * 16)
    # preconditioned GMRES solves the system (true residual) and beats the one-level preconditioner
    dx1, its1, ok1 = eng.linear_solve(F, 1e-11, 1e-50, 200, 2000)
    dx1 = dx1.clone()
    dx2, its2, ok2 = eng.linear_solve(F, 1e-11, 1e-50, 200, 2000, mg=True)
    assert ok1 and ok2
    b = F.cpu().numpy()
    for dx in (dx1, dx2):
        assert np.linalg.norm(A @ dx.cpu().numpy() - b) <= 1e-9 * np.linalg.norm(b)
    assert np.abs((dx1 - dx2).cpu().numpy()).max() <= 1e-6 * np.abs(dx1.cpu().numpy()).max()
    assert its2 < its1 and its2 <= 40, (its1, its2)


def test_gmg_option_set_newton_solution():
    """init_solver_parameters('gmg'): FGMRES(200) rtol 1e-2 around the V-cycle; the Newton solution is the MMS one."""
    from echemfem_b200.fd import errornorm
    s = product_cases.AdvectionDiffusionMigrationSolver(32, 1.0, D1=0.5e-5, D2=1.0e-5, mesh=UnitSquareMesh(32, 32, tile=8))
    s.init_solver_parameters(pc_type="gmg")
    s.solver_parameters.pop("snes_monitor", None)
    s.solver_parameters.pop("ksp_monitor", None)
    s.setup_solver()
    s.solve()
    c1, U = s.u.subfunctions
    assert errornorm(s.C1ex, c1) < 2e-4 and errornorm(s.Uex, U) < 2e-4
    e = s._engine
    assert e.multigrid() is not None and max(e.last["ksp_its_per_step"]) <= 30


def test_multigrid_on_extruded_hexes():
    """3D (DQ1 x DG1 hexes, the configs[3] element): the tensor hierarchy, prolongation and V-cycle in three dimensions."""
    from echemfem_b200.mesh import ExtrudedMesh
    from echemfem_b200.newton import NewtonEngine
    from echemfem_b200.multigrid import Multigrid
    mesh = ExtrudedMesh(UnitSquareMesh(8, 8), 8, 1.0 / 8)
    s = product_cases.AdvectionDiffusionMigrationSolver(0, 1.0, mesh=mesh)
    eng = NewtonEngine(s, nblocks=64)
    x = s.V.node_coords()
    u = np.concatenate([mms_cases.C1ex(x), mms_cases.Uex(x) + 1e-3 * np.sin(7 * x[:, 0] + 3 * x[:, 1] + x[:, 2])])
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F = eng.residual(s.u.tensor).clone()
    eng.jacobian(s.u.tensor)
    eng._mg = Multigrid(eng, tile=4, cells_per_block=8, coarsest_cells=64)
    assert eng._mg.nlevels == 2 and eng._mg.h.levels[1].shape == (4, 4, 4)
    rp, ci, v = eng.csr_host()
    A = sp.csr_matrix((v, ci, rp), shape=(eng.ndof, eng.ndof))
    dx1, its1, ok1 = eng.linear_solve(F, 1e-10, 1e-50, 200, 2000)
    dx1 = dx1.clone()
    dx2, its2, ok2 = eng.linear_solve(F, 1e-10, 1e-50, 200, 2000, mg=True)
    assert ok1 and ok2 and its2 < its1
    b = F.cpu().numpy()
    assert np.linalg.norm(A @ dx2.cpu().numpy() - b) <= 1e-8 * np.linalg.norm(b)


@pytest.mark.parametrize("tile", [None, 8])
def test_auxiliary_space_operator_and_iteration_count(tile, monkeypatch):
    """The conforming auxiliary space of the potential (auxspace.py): A_aux = Q^T A_phiphi Q and its coarse levels
    against scipy, and the V-cycle with it needs fewer GMRES iterations than the nested DG hierarchy alone."""
    from echemfem_b200.multigrid import Multigrid
    eng, A, F = _system(64, tile)
    M = Multigrid(eng, tile=4, cells_per_block=16, coarsest_cells=16)
    aux = M.aux
    assert aux is not None and aux.fields == [1]
    M.setup()
    N = eng.N
    Q = sp.csr_matrix(tuple(t.cpu().numpy() for t in aux.Q)[::-1], shape=(N, aux.ncg))
    App = A[N:2 * N][:, N:2 * N]
    want = (Q.T @ App @ Q).tocsr()
    L0 = aux.lv[0]
    got = sp.csr_matrix((L0.vals.cpu().numpy(), L0.colidx.cpu().numpy(), L0.rowptr.cpu().numpy()), shape=want.shape)
    assert abs(got - want).max() <= 1e-12 * abs(want).max()
    # a continuous function has no jumps: the embedding of a linear vertex function is linear at the DG nodes
    x = eng.solver.V.node_coords()
    g = np.meshgrid(np.linspace(0, 1, 65), np.linspace(0, 1, 65), indexing="ij")
    np.testing.assert_allclose(Q @ (2 * g[0].ravel() - 3 * g[1].ravel()), 2 * x[:, 0] - 3 * x[:, 1], atol=1e-13)
    # next conforming level: P^T A P with the host interpolation
    P = sp.csr_matrix(tuple(t.cpu().numpy() for t in L0.P)[::-1], shape=(L0.nnode, aux.lv[1].nnode))
    L1 = aux.lv[1]
    got1 = sp.csr_matrix((L1.vals.cpu().numpy(), L1.colidx.cpu().numpy(), L1.rowptr.cpu().numpy()), shape=(L1.n, L1.n))
    want1 = (P.T @ want @ P).tocsr()
    assert abs(got1 - want1).max() <= 1e-11 * abs(want1).max()
    # iteration counts: with / without the auxiliary correction, same solution
    b = F.cpu().numpy()
    dx2, its2, ok2 = eng.linear_solve(F, 1e-11, 1e-50, 200, 2000, mg=True)
    dx2 = dx2.clone()
    monkeypatch.setenv("ECH_MG_AUX", "0")
    del eng._mg
    dx1, its1, ok1 = eng.linear_solve(F, 1e-11, 1e-50, 200, 2000, mg=True)
    assert eng._mg.aux is None
    assert ok1 and ok2
    for dx in (dx1, dx2):
        assert np.linalg.norm(A @ dx.cpu().numpy() - b) <= 1e-9 * np.linalg.norm(b)
    assert its2 < its1, (its1, its2)
    print("GMRES iterations N=64: nested DG hierarchy %d, with the conforming auxiliary space %d" % (its1, its2))


def test_global_conforming_coarse_operator():
    """E^T A E of the coarse space shared by all ranks (auxspace.GlobalConformingCoarse): the kernel's ELL result on the
    CSR pattern of level 0 against scipy, on one rank (no ghosts); the multi-rank sum is covered by the 2-rank run of
    tests/test_gpu_distributed.py."""
    from echemfem_b200.auxspace import GlobalConformingCoarse
    eng, A, F = _system(32, 8)
    axes = [np.linspace(0.0, 1.0, 33), np.linspace(0.0, 1.0, 33)]
    G = GlobalConformingCoarse(eng, [1], axes, level=2)
    assert G.level == 2 and G.vdims == [9, 9]
    vals = G.local_operator(eng._stream()).cpu().numpy()
    L0 = G.h.lv[0]
    got = sp.csr_matrix((vals, L0.colidx.cpu().numpy(), L0.rowptr.cpu().numpy()), shape=(81, 81))
    N = eng.N
    E = sp.csr_matrix(tuple(t.cpu().numpy() for t in G.E)[::-1], shape=(N, 81))
    want = (E.T @ A[N:2 * N][:, N:2 * N] @ E).tocsr()
    assert abs(got - want).max() <= 1e-12 * abs(want).max()
    # E evaluates coarse vertex functions at the DG nodes: linear functions are reproduced
    x = eng.solver.V.node_coords()
    g = np.meshgrid(axes[0][::4], axes[1][::4], indexing="ij")
    np.testing.assert_allclose(E @ (g[0].ravel() - 2 * g[1].ravel()), x[:, 0] - 2 * x[:, 1], atol=1e-13)
