"""Two-level preconditioner (aggregation coarse space inside the native GMRES): same solution as a
direct solve of the assembled Jacobian, fewer Krylov iterations than block-Jacobi/ILU(0) alone."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import mms_cases
import product_cases

pytestmark = pytest.mark.gpu


def _system(N):
    from echemfem_b200.newton import NewtonEngine
    s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0)
    eng = NewtonEngine(s, nblocks=16)
    x = s.V.node_coords()
    u = np.concatenate([mms_cases.C1ex(x), mms_cases.Uex(x) + 1e-3 * np.sin(7 * x[:, 0] + 3 * x[:, 1])])
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F = eng.residual(s.u.tensor).clone()
    eng.jacobian(s.u.tensor)
    rp, ci, v = eng.csr_host()
    A = sp.csr_matrix((v, ci, rp), shape=(eng.ndof, eng.ndof))
    return eng, A, F


@pytest.mark.parametrize("mode", [2, 3])
def test_two_level_gmres_matches_direct_solve(mode):
    eng, A, F = _system(24)
    ref = spla.spsolve(A.tocsc(), F.cpu().numpy())
    dx1, its1, ok1 = eng.linear_solve(F, 1e-12, 1e-50, 200, 5000)
    dx1 = dx1.clone()
    dx2, its2, ok2 = eng.linear_solve(F, 1e-12, 1e-50, 200, 5000, coarse=128, coarse_mode=mode)
    assert ok1 and ok2
    scale = np.abs(ref).max()
    assert np.abs(dx1.cpu().numpy() - ref).max() <= 1e-8 * scale
    assert np.abs(dx2.cpu().numpy() - ref).max() <= 1e-8 * scale
    assert its2 < its1, "coarse level did not reduce the iteration count (%d vs %d)" % (its2, its1)


def test_coarse_kernels_match_dense_algebra():
    """P^T A P, P^T v and z += P y from the CUDA kernels against scipy on the same aggregation."""
    from echemfem_b200 import capi
    from echemfem_b200.capi import ptr
    eng, A, F = _system(8)
    Cz = eng._coarse_plan(32)
    nc, agg = Cz["nc"], Cz["agg"].cpu().numpy()
    P = sp.csr_matrix((np.ones(eng.ndof), (np.arange(eng.ndof), agg)), shape=(eng.ndof, nc))
    st = eng._stream()
    capi.check(eng.lib.ech_coarse_galerkin(eng.ndof, ptr(eng.rowptr), ptr(eng.colidx), ptr(eng.vals), ptr(Cz["agg"]),
                                           nc, ptr(Cz["Ac"]), st))
    Ac = (P.T @ A @ P).toarray()
    assert np.abs(Cz["Ac"].cpu().numpy() - Ac).max() <= 1e-12 * np.abs(Ac).max()
    rc = torch.empty(nc, dtype=torch.float64, device="cuda")
    capi.check(eng.lib.ech_coarse_restrict(eng.ndof, ptr(Cz["agg"]), nc, ptr(F), ptr(rc), st))
    want = P.T @ F.cpu().numpy()
    assert np.abs(rc.cpu().numpy() - want).max() <= 1e-12 * np.abs(want).max()
    y = torch.from_numpy(np.cos(np.arange(nc) * 0.3)).cuda()
    z = F.clone()
    capi.check(eng.lib.ech_coarse_prolong_add(eng.ndof, ptr(Cz["agg"]), ptr(y), ptr(z), st))
    want = F.cpu().numpy() + P @ y.cpu().numpy()
    assert np.abs(z.cpu().numpy() - want).max() <= 1e-14 * np.abs(want).max()
    M = torch.from_numpy(np.sin(np.arange(nc * nc, dtype=float)).reshape(nc, nc)).cuda()
    out = torch.empty(nc, dtype=torch.float64, device="cuda")
    capi.check(eng.lib.ech_dense_gemv(nc, nc, ptr(M), ptr(y), ptr(out), st))
    want = M.cpu().numpy() @ y.cpu().numpy()
    assert np.abs(out.cpu().numpy() - want).max() <= 1e-12 * np.abs(want).max()
