"""GPU tests of the host-buffer boundary `ech_assemble_host` (include/echem_b200.h): the call a binding inside the
reference's SNES callbacks would make (`NonlinearVariationalSolver.solve`, echemfem/solver.py:714-725) with the
state in host memory.  The chunked three-stream pipeline must give bit-identical results to the one-launch path,
whatever the chunk count and the cell numbering."""
import numpy as np
import pytest
import torch

import product_cases
from echemfem_b200 import capi
from echemfem_b200.capi import ptr
from echemfem_b200.mesh import UnitSquareMesh, ExtrudedMesh

pytestmark = pytest.mark.gpu


def _setup(mesh, **kw):
    from echemfem_b200.newton import NewtonEngine
    s = product_cases.AdvectionDiffusionMigrationSolver(0, 3.0, mesh=mesh, **kw)
    e = NewtonEngine(s)
    x = s.V.node_coords()
    z = x[:, 2] if x.shape[1] > 2 else 0.0
    u = np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3 + 1e-2 * np.sin(9 * x[:, 0] * x[:, 1] + z),
                        np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3 + 0.1 * z])
    return s, e, u


def _host_call(e, s, u_host, chunks):
    lib = e.lib
    capi.check(lib.ech_set_host_chunks(e.h, chunks), e.h)
    uh = torch.from_numpy(u_host.copy()).pin_memory()
    Fh = torch.full((u_host.size,), np.nan, dtype=torch.float64).pin_memory()
    u_dev = torch.full_like(s.u.tensor, np.nan)
    e.vals.fill_(float("nan"))
    e.F.fill_(float("nan"))
    aux, consts = e._aux_tensor(), e._consts()      # held: the call takes raw pointers
    n0 = lib.ech_launch_count(e.h)
    capi.check(lib.ech_assemble_host(e.h, ptr(uh), ptr(u_dev), ptr(aux), ptr(consts), ptr(e.rowptr),
                                     ptr(e.vals), ptr(e.F), ptr(Fh), e._stream()), e.h)
    launches = lib.ech_launch_count(e.h) - n0
    assert torch.equal(u_dev.cpu(), uh)
    return Fh.numpy().copy(), e.vals.cpu().numpy().copy(), launches


@pytest.mark.parametrize("tile", [None, 4])
@pytest.mark.parametrize("N", [24, 29])
def test_pipelined_host_assembly_is_bit_identical(N, tile):
    mesh = UnitSquareMesh(N, N, tile=tile) if tile else UnitSquareMesh(N, N)
    s, e, u = _setup(mesh)
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F_dev = e.residual_jacobian(s.u.tensor)[0].cpu().numpy().copy()
    J_dev = e.vals.cpu().numpy().copy()
    assert np.isfinite(F_dev).all() and np.isfinite(J_dev).all()
    F1, J1, l1 = _host_call(e, s, u, 0)
    assert l1 == 1
    assert np.array_equal(F1, F_dev) and np.array_equal(J1, J_dev)
    for chunks in (2, 3, 8, 13):
        Fk, Jk, lk = _host_call(e, s, u, chunks)
        assert 2 <= lk <= chunks
        assert np.array_equal(Fk, F_dev), chunks
        assert np.array_equal(Jk, J_dev), chunks
    # repeated calls reuse the pipeline
    Fk, Jk, _ = _host_call(e, s, u + 0.0, 13)
    assert np.array_equal(Fk, F_dev) and np.array_equal(Jk, J_dev)


def test_pipelined_host_assembly_3d():
    mesh = ExtrudedMesh(UnitSquareMesh(6, 5), 4, 0.25)
    s, e, u = _setup(mesh)
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F_dev = e.residual_jacobian(s.u.tensor)[0].cpu().numpy().copy()
    J_dev = e.vals.cpu().numpy().copy()
    for chunks in (0, 4, 7):
        Fk, Jk, _ = _host_call(e, s, u, chunks)
        assert np.array_equal(Fk, F_dev) and np.array_equal(Jk, J_dev), chunks


def test_host_assembly_falls_back_to_one_launch_where_ranges_are_not_supported():
    """Q3 (first-generation kernel) and CG modules have no cell-range launch: one copy, one call, one copy."""
    s, e, u = _setup(UnitSquareMesh(12, 12), p=3)
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F_dev = e.residual_jacobian(s.u.tensor)[0].cpu().numpy().copy()
    J_dev = e.vals.cpu().numpy().copy()
    Fk, Jk, lk = _host_call(e, s, u, 8)
    assert np.array_equal(Fk, F_dev) and np.array_equal(Jk, J_dev)
