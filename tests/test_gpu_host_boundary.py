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


def _host_call(e, s, u_host, chunks, with_matrix=True):
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
                                     ptr(e.vals) if with_matrix else None, ptr(e.F), ptr(Fh), e._stream()), e.h)
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


def test_pipelined_host_assembly_q3_first_generation_kernel():
    """Q3 runs the first-generation cooperative Jacobian kernel; it takes cell ranges too."""
    s, e, u = _setup(UnitSquareMesh(12, 12), p=3)
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F_dev = e.residual_jacobian(s.u.tensor)[0].cpu().numpy().copy()
    J_dev = e.vals.cpu().numpy().copy()
    for chunks in (0, 5, 8):
        Fk, Jk, lk = _host_call(e, s, u, chunks)
        assert (lk == 1) if chunks < 2 else (2 <= lk <= chunks)
        assert np.array_equal(Fk, F_dev) and np.array_equal(Jk, J_dev), chunks


@pytest.mark.parametrize("p", [1, 2])
def test_pipelined_host_residual_only(p):
    """The SNES function callback alone (line search): vals = NULL, residual kernel per cell range."""
    s, e, u = _setup(UnitSquareMesh(20, 17), p=p)
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F_dev = e.residual(s.u.tensor).cpu().numpy().copy()
    for chunks in (0, 2, 7, 16):
        Fk, Jk, lk = _host_call(e, s, u, chunks, with_matrix=False)
        assert (lk == 1) if chunks < 2 else (2 <= lk <= chunks)
        assert np.isnan(Jk).all()                   # the matrix was not touched
        assert np.array_equal(Fk, F_dev), chunks


def test_host_assembly_of_a_continuous_space_is_one_launch_set():
    """CG modules have no cell-range launch (rows are shared between cells): one copy, one call, one copy."""
    from echemfem_b200.newton import NewtonEngine
    s = product_cases.AdvectionDiffusionMigrationSolver(12, 1.0, family="CG")
    e = NewtonEngine(s)
    rng = np.random.default_rng(3)
    u = 3.0 + 0.1 * rng.standard_normal(s.u.tensor.numel())
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F_dev = e.residual_jacobian(s.u.tensor)[0].cpu().numpy().copy()
    J_dev = e.vals.cpu().numpy().copy()
    Fk, Jk, _ = _host_call(e, s, u, 8)
    assert np.array_equal(Fk, F_dev) and np.array_equal(Jk, J_dev)


@pytest.mark.parametrize("size,how", [(2, "ranges"), (3, "graph")])
def test_pipelined_host_assembly_with_ghost_cells(size, how):
    """A rank's local mesh (owned + ghost cells): ghost state travels ahead of the first range, ghost
    rows ride with the last one; results equal the one-launch path bit for bit."""
    from echemfem_b200.newton import NewtonEngine
    from echemfem_b200.partition import partition_mesh, graph_partition
    mesh = UnitSquareMesh(20, 20)
    parts = graph_partition(mesh, size) if how == "graph" else None
    rng = np.random.default_rng(11)
    for rank in range(size):
        local, plan = partition_mesh(mesh, rank, size, parts=parts)
        assert local.num_cells > local.num_owned
        s = product_cases.AdvectionDiffusionMigrationSolver(0, 3.0, mesh=local)
        e = NewtonEngine(s)
        e.halo = None                                        # ghost values are part of the host state here
        u = 3.0 + 0.1 * rng.standard_normal(s.u.tensor.numel())
        s.u.tensor.copy_(torch.from_numpy(u).cuda())
        e.F.zero_()
        F_dev = e.residual_jacobian(s.u.tensor)[0].cpu().numpy().copy()
        J_dev = e.vals.cpu().numpy().copy()
        n = 4
        owned = np.concatenate([f * local.num_cells * n + np.arange(local.num_owned * n) for f in range(2)])
        for chunks in (0, 4, 9):
            Fk, Jk, _ = _host_call(e, s, u, chunks)
            assert np.array_equal(Fk[owned], F_dev[owned]), (rank, chunks)
            assert np.array_equal(Jk, J_dev), (rank, chunks)


def test_auto_mode_decides_between_pipeline_and_single_copy():
    """`ech_set_host_chunks(h, -k)`: the first six calls time the k-chunk pipeline and the single-copy path, every one
    of them a valid call (bit-identical results); afterwards the faster mode stays selected."""
    s, e, u = _setup(UnitSquareMesh(40, 40))
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F_dev = e.residual_jacobian(s.u.tensor)[0].cpu().numpy().copy()
    J_dev = e.vals.cpu().numpy().copy()
    lib = e.lib
    capi.check(lib.ech_set_host_chunks(e.h, -4), e.h)
    uh = torch.from_numpy(u.copy()).pin_memory()
    Fh = torch.empty(u.size, dtype=torch.float64).pin_memory()
    aux, consts = e._aux_tensor(), e._consts()
    launches = []
    for k in range(9):
        Fh.fill_(float("nan"))
        e.vals.fill_(float("nan"))
        n0 = lib.ech_launch_count(e.h)
        capi.check(lib.ech_assemble_host(e.h, ptr(uh), ptr(s.u.tensor), ptr(aux), ptr(consts), ptr(e.rowptr),
                                         ptr(e.vals), ptr(e.F), ptr(Fh), e._stream()), e.h)
        launches.append(lib.ech_launch_count(e.h) - n0)
        assert np.array_equal(Fh.numpy(), F_dev) and np.array_equal(e.vals.cpu().numpy(), J_dev), k
    assert all(2 <= n <= 4 for n in launches[:3]) and launches[3:6] == [1, 1, 1]        # measured both modes
    assert len(set(launches[6:])) == 1                                                   # then one mode, kept
