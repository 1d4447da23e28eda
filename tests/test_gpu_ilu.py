"""Block-Jacobi / ILU(0) on the device (`ech_bjacobi_ilu0_factor / _apply`, the `ilu` option set of
echemfem/solver.py:938-957: PETSc bjacobi + ilu(0) per block) against a host ILU(0) of every diagonal block."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import product_cases
from echemfem_b200 import capi
from echemfem_b200.capi import ptr
from echemfem_b200.mesh import UnitSquareMesh, ExtrudedMesh

pytestmark = pytest.mark.gpu


def _host_block_ilu0_solve(A, rows, r):
    """ILU(0) of A[rows][:, rows] on its own pattern (IKJ), then L U z = r[rows]; rows ascending."""
    S = sp.csr_matrix(A[rows][:, rows])
    S.sort_indices()
    n = S.shape[0]
    ip, ix, v = S.indptr, S.indices, S.data.copy()
    pos = [dict(zip(ix[ip[i]:ip[i + 1]], range(ip[i], ip[i + 1]))) for i in range(n)]
    for i in range(n):
        for p in range(ip[i], ip[i + 1]):
            k = ix[p]
            if k >= i:
                break
            v[p] /= v[pos[k][k]]
            for q in range(ip[k], ip[k + 1]):
                j = ix[q]
                if j > k and j in pos[i]:
                    v[pos[i][j]] -= v[p] * v[q]
    z = r[rows].copy()
    for i in range(n):
        for p in range(ip[i], ip[i + 1]):
            if ix[p] < i:
                z[i] -= v[p] * z[ix[p]]
    for i in range(n - 1, -1, -1):
        for p in range(ip[i], ip[i + 1]):
            if ix[p] > i:
                z[i] -= v[p] * z[ix[p]]
        z[i] /= v[pos[i][i]]
    return z


@pytest.mark.parametrize("case", ["q1", "q2", "q3", "hex"])
def test_block_ilu0_apply_matches_host(case):
    from echemfem_b200.newton import NewtonEngine
    if case == "hex":
        mesh, p, cells_per_block = ExtrudedMesh(UnitSquareMesh(4, 3), 3, 0.3), 1, 7
    elif case == "q3":
        mesh, p, cells_per_block = UnitSquareMesh(4, 3), 3, 5            # half-rows longer than 64 entries
    else:
        mesh, p, cells_per_block = UnitSquareMesh(9, 7), (1 if case == "q1" else 2), 10
    s = product_cases.AdvectionDiffusionMigrationSolver(0, 3.0, mesh=mesh, p=p)
    e = NewtonEngine(s)
    x = s.V.node_coords()
    u = np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3, np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3])
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    e.jacobian(s.u.tensor)
    rp, ci, v = e.csr_host()
    # the electroneutral Jacobian has negligible ILU pivots (the device shifts them, PETSc shift_type nonzero):
    # compare on a diagonally dominant matrix with the same pattern
    rows_of = np.repeat(np.arange(e.ndof), np.diff(rp))
    dpos = np.nonzero(ci == rows_of)[0]
    assert dpos.size == e.ndof
    v = v.copy()
    v[dpos] += 2.0 * np.bincount(rows_of, weights=np.abs(v), minlength=e.ndof)
    e.vals.copy_(torch.from_numpy(v).cuda())
    A = sp.csr_matrix((v, ci, rp), shape=(e.ndof, e.ndof))
    ncell, n, nf = mesh.num_cells, e.nloc, e.nf
    bounds = list(range(0, ncell, cells_per_block)) + [ncell]            # ragged last block
    block_ptr = torch.tensor(bounds, dtype=torch.int64, device="cuda")
    nblocks = len(bounds) - 1
    lu = torch.empty_like(e.vals)
    diag = torch.empty(e.ndof, dtype=torch.int32, device="cuda")
    st = e._stream()
    capi.check(e.lib.ech_bjacobi_ilu0_factor(e.ndof, nf, ncell * n, n, nblocks, ptr(block_ptr), ptr(e.rowptr),
                                             ptr(e.colidx), ptr(e.vals), ptr(lu), ptr(diag), st))
    rng = np.random.default_rng(5)
    r = rng.standard_normal(e.ndof)
    rd = torch.from_numpy(r).cuda()
    z = torch.full_like(rd, float("nan"))
    capi.check(e.lib.ech_bjacobi_ilu0_apply(e.ndof, nf, ncell * n, n, nblocks, ptr(block_ptr), ptr(e.rowptr),
                                            ptr(e.colidx), ptr(lu), ptr(diag), ptr(rd), ptr(z), st))
    z = z.cpu().numpy()
    assert np.isfinite(z).all()
    # level-scheduled solves (ech_bjacobi_ilu0_plan_* / _apply_planned) on the same factors, every lane grouping
    from echemfem_b200.ilu import IluPlan
    zp = {}
    for lpr in (None, 2, 16):
        plan = IluPlan(e.lib, e.dev, st, e.ndof, nf, ncell * n, n, np.array(bounds), block_ptr, e.rowptr, e.colidx,
                       e.nnz)
        assert plan.ok
        if lpr is not None and lpr != plan.lpr:          # rebuild with another grouping: same result up to rounding
            plan.lpr = lpr
            plan.packed = None
            G = 32 // lpr
            levels = torch.empty(2 * e.ndof, dtype=torch.int16, device="cuda")
            import ctypes as C
            npass = C.c_int64()
            capi.check(e.lib.ech_bjacobi_ilu0_plan_count(e.ndof, nf, ncell * n, n, nblocks, ptr(block_ptr), plan.max_rows,
                                                         ptr(e.rowptr), ptr(e.colidx), lpr, ptr(plan.lcol), ptr(levels),
                                                         ptr(plan.passptr), C.byref(npass), st))
            plan.npass = int(npass.value)
            plan.slots = torch.empty(max(npass.value * G, 1), dtype=torch.int64, device="cuda")
            capi.check(e.lib.ech_bjacobi_ilu0_plan_fill(e.ndof, nf, ncell * n, n, nblocks, ptr(block_ptr), plan.max_rows,
                                                        ptr(e.rowptr), ptr(e.colidx), lpr, ptr(levels),
                                                        ptr(plan.passptr), ptr(plan.slots), st))
        zz = torch.full_like(rd, float("nan"))
        if lpr is None:
            # the factorisation on the plan's local columns gives the factors of the search-based kernel (same
            # arithmetic in the same order; a ragged last block and blocks of any field count included)
            lu2 = torch.full_like(lu, float("nan"))
            diag2 = torch.empty_like(diag)
            plan.factor(e.rowptr, e.colidx, e.vals, lu2, diag2, st)
            assert torch.equal(diag, diag2)
            a, b2 = lu.cpu().numpy(), lu2.cpu().numpy()
            assert np.isfinite(b2).all()
            assert np.abs(a - b2).max() <= 1e-14 * np.abs(a).max(), (case, np.abs(a - b2).max())
            # pass-packed factors fetched with cp.async.bulk (the default path) ...
            assert plan.packed is not None
            plan.refresh(lu, st)
            plan.apply(lu, rd, zz, st)
            zp["packed"] = zz.cpu().numpy()
            assert np.isfinite(zp["packed"]).all(), case
            plan._packed_for = None                       # ... and the register-pipelined planned solve
            zz = torch.full_like(rd, float("nan"))
        plan.apply(lu, rd, zz, st)
        zp[plan.lpr] = zz.cpu().numpy()
        assert np.isfinite(zp[plan.lpr]).all(), (case, plan.lpr)
        # every row has a slot in each sweep; a pass holds up to 32 / lpr rows of one dependency level
        assert plan.npass * (32 // plan.lpr) >= 2 * e.ndof and plan.npass <= 2 * e.ndof
    for b in range(nblocks):
        c0, c1 = bounds[b], bounds[b + 1]
        rows = np.concatenate([f * ncell * n + np.arange(c0 * n, c1 * n) for f in range(nf)])
        zh = _host_block_ilu0_solve(A, rows, r)
        assert np.abs(z[rows] - zh).max() <= 1e-10 * max(np.abs(zh).max(), 1.0), (case, b)
        for lpr, zz in zp.items():
            assert np.abs(zz[rows] - zh).max() <= 1e-10 * max(np.abs(zh).max(), 1.0), (case, b, lpr)
