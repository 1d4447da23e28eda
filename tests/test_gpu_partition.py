"""Ghost-cell layout on the device: each rank's local assembly equals its rows of the global assembly.

Ranks are emulated one after the other on a single GPU (ghost values filled from the global state);
the exchange itself is covered by tests/test_partition.py (gloo) and bench.py --gpus N (NCCL).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import product_cases
from echemfem_b200.mesh import UnitSquareMesh
from echemfem_b200.partition import partition_mesh

pytestmark = pytest.mark.gpu


def _global_csr(eng):
    rp, ci, v = eng.csr_host()
    return sp.csr_matrix((v, ci, rp), shape=(eng.ndof, eng.ndof))


@pytest.mark.parametrize("size,how", [(2, "ranges"), (3, "ranges"), (4, "graph")])
def test_local_assembly_equals_global_rows(size, how):
    from echemfem_b200.newton import NewtonEngine
    from echemfem_b200.partition import graph_partition
    N, n, nf = 6, 4, 2
    mesh = UnitSquareMesh(N, N)
    sg = product_cases.AdvectionDiffusionMigrationSolver(0, 3.0, mesh=mesh)
    eg = NewtonEngine(sg)
    x = sg.V.node_coords()
    ug = np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3, np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3])
    ug = ug + 1e-2 * np.sin(5 * np.arange(len(ug)))
    sg.u.tensor.copy_(torch.from_numpy(ug).cuda())
    Fg = eg.residual(sg.u.tensor).cpu().numpy()
    eg.jacobian(sg.u.tensor)
    Jg = _global_csr(eg)
    Ng = mesh.num_cells * n
    parts = graph_partition(mesh, size) if how == "graph" else None
    for rank in range(size):
        local, plan = partition_mesh(mesh, rank, size, parts=parts)
        s = product_cases.AdvectionDiffusionMigrationSolver(0, 3.0, mesh=local)
        e = NewtonEngine(s)
        e.halo = None                                        # ghosts are filled by hand below
        gid = local.global_ids
        Nl = local.num_cells * n
        gdof = (gid[:, None] * n + np.arange(n)[None, :]).reshape(-1)            # local scalar dof -> global
        ul = np.concatenate([ug[f * Ng + gdof] for f in range(nf)])
        s.u.tensor.copy_(torch.from_numpy(ul).cuda())
        Fl = e.residual(s.u.tensor).cpu().numpy()
        e.jacobian(s.u.tensor)
        Jl = _global_csr(e)
        no = local.num_owned * n
        gmap = np.concatenate([f * Ng + gdof for f in range(nf)])                 # local dof -> global dof
        owned_rows = np.concatenate([f * Nl + np.arange(no) for f in range(nf)])
        scale = np.abs(Fg).max()
        assert np.abs(Fl[owned_rows] - Fg[gmap[owned_rows]]).max() <= 1e-12 * scale
        # ghost rows are empty
        ghost_rows = np.setdiff1d(np.arange(nf * Nl), owned_rows)
        assert Jl[ghost_rows].nnz == 0 and np.all(Fl[ghost_rows] == 0)
        # owned rows, mapped to global numbering, equal the global rows
        P = sp.csr_matrix((np.ones(nf * Nl), (np.arange(nf * Nl), gmap)), shape=(nf * Nl, nf * Ng))
        Jl_glob = (Jl @ P)[owned_rows]
        D = Jl_glob - Jg[gmap[owned_rows]]
        assert abs(D).max() <= 1e-12 * abs(Jg).max()
        assert Jl[owned_rows].nnz == Jg[gmap[owned_rows]].nnz


def test_tile_numbered_mesh_gives_the_same_operator():
    """Cells numbered tile by tile (mesh.tile_order): residual and Jacobian equal those of the lexicographic mesh
    after permuting the dofs -- numbering is not part of parity but must not change the operator."""
    from echemfem_b200.newton import NewtonEngine
    N, n, nf = 8, 4, 2
    out = {}
    for tile in (None, 4):
        mesh = UnitSquareMesh(N, N, tile=tile)
        s = product_cases.AdvectionDiffusionMigrationSolver(0, 3.0, mesh=mesh)
        e = NewtonEngine(s)
        x = s.V.node_coords()
        u = np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3 + 1e-2 * np.sin(9 * x[:, 0] * x[:, 1]),
                            np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3])
        s.u.tensor.copy_(torch.from_numpy(u).cuda())
        F = e.residual(s.u.tensor).cpu().numpy()
        e.jacobian(s.u.tensor)
        out[tile] = (mesh, F, _global_csr(e))
    mesh_t, Ft, Jt = out[4]
    _, Fl, Jl = out[None]
    perm = mesh_t.cell_perm                                   # tile cell i is lexicographic cell perm[i]
    Ns = N * N * n
    dof = (perm[:, None] * n + np.arange(n)[None, :]).reshape(-1)
    gmap = np.concatenate([f * Ns + dof for f in range(nf)])    # tile dof -> lexicographic dof
    assert np.abs(Ft - Fl[gmap]).max() <= 1e-12 * np.abs(Fl).max()
    D = Jt - Jl[gmap][:, gmap]
    assert abs(D).max() <= 1e-12 * abs(Jl).max() and Jt.nnz == Jl.nnz
