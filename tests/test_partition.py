"""Partition + halo exchange host logic; the N > 1 path is exercised with world_size-2 gloo on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from echemfem_b200.mesh import TensorMesh, UnitSquareMesh
from echemfem_b200.partition import partition_mesh, strip_partition, HaloExchange


def _check_local(mesh, local):
    """Local connectivity, mapped back to global ids, is the global connectivity of the owned cells."""
    gid = local.global_ids
    no = local.num_owned
    for lc in range(no):
        g = gid[lc]
        for lf in range(2 * mesh.dim):
            ln = local.nbr_cell[lc, lf]
            gn = mesh.nbr_cell[g, lf]
            assert (ln < 0) == (gn < 0)
            if ln >= 0:
                assert gid[ln] == gn
            else:
                assert local.ext_marker[lc, lf] == mesh.ext_marker[g, lf]
    assert np.allclose(local.cell_coords(), mesh.cell_coords()[gid])


@pytest.mark.parametrize("size", [1, 2, 3])
def test_partition_covers_mesh(size):
    mesh = UnitSquareMesh(6, 4)
    owned = []
    for r in range(size):
        local, plan = partition_mesh(mesh, r, size)
        _check_local(mesh, local)
        owned += list(local.global_ids[:local.num_owned])
        for peer, cells in plan.recv.items():
            assert np.all(cells >= local.num_owned)
        for peer, cells in plan.send.items():
            assert np.all(cells < local.num_owned)
    assert sorted(owned) == list(range(mesh.num_cells))


def test_strip_partition_equals_global_partition():
    ax = [np.linspace(0, 3, 10), np.linspace(0, 1, 5)]
    mesh = TensorMesh(ax)
    for size in (2, 3):
        for r in range(size):
            local, plan = strip_partition(ax, r, size)
            edges = np.linspace(0, 9, size + 1).astype(np.int64) * 4
            ref, rplan = partition_mesh(mesh, r, size, edges=edges)
            assert np.array_equal(local.global_ids, ref.global_ids)
            assert np.array_equal(local.nbr_cell, ref.nbr_cell)
            assert np.array_equal(local.ext_marker, ref.ext_marker)
            for peer in rplan.peers():
                assert np.array_equal(plan.send[peer], rplan.send[peer])
                assert np.array_equal(plan.recv[peer], rplan.recv[peer])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, ok, graph=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        mesh = UnitSquareMesh(6, 4)
        parts = None
        if graph:
            from echemfem_b200.partition import graph_partition
            parts = graph_partition(mesh, size)
        local, plan = partition_mesh(mesh, rank, size, parts=parts)
        nf, nloc = 2, 4
        ntot = local.num_cells
        hx = HaloExchange(plan, nf, nloc, ntot, torch.device("cpu"))
        gid = torch.from_numpy(local.global_ids)
        # value of dof (field f, global cell g, node b)
        val = (torch.arange(nf)[:, None, None] * 1000.0 + gid[None, :, None] * 10.0 + torch.arange(nloc)[None, None, :])
        u = val.clone().double()
        u[:, local.num_owned:, :] = -1.0                         # ghosts unknown before the exchange
        u = u.reshape(-1).contiguous()
        hx.exchange(u)
        assert torch.equal(u, val.double().reshape(-1))
        # allreduce as used for Krylov dots
        t = torch.tensor([float(local.num_owned)], dtype=torch.float64)
        dist.all_reduce(t)
        assert int(t.item()) == mesh.num_cells
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("graph", [False, True])
def test_halo_exchange_gloo_world2(graph):
    size = 2
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * size)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, ok, graph)) for r in range(size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert list(ok) == [1] * size


@pytest.mark.parametrize("size", [2, 3, 8])
def test_graph_partition_is_balanced_connected_and_consistent(size):
    """Graph partitioner on a graded (Bortels) mesh: every cell owned once, parts balanced to one cell, each
    part connected, cut small compared with a random assignment; local meshes built from it are consistent."""
    from scipy.sparse.csgraph import connected_components
    import scipy.sparse as sp
    from echemfem_b200.meshes import BortelsMesh
    from echemfem_b200.partition import graph_partition
    mesh = BortelsMesh(9, 7, 9, 8)
    parts = graph_partition(mesh, size)
    counts = np.bincount(parts, minlength=size)
    assert counts.sum() == mesh.num_cells and counts.max() - counts.min() <= 1
    nb = mesh.nbr_cell
    rows = np.repeat(np.arange(mesh.num_cells), nb.shape[1])
    cols = nb.reshape(-1)
    keep = cols >= 0
    cut = int((parts[rows[keep]] != parts[cols[keep]]).sum()) // 2
    rnd = np.random.default_rng(0).permutation(parts)
    cut_rnd = int((rnd[rows[keep]] != rnd[cols[keep]]).sum()) // 2
    assert cut < 0.35 * cut_rnd
    G = sp.csr_matrix((np.ones(int(keep.sum())), (rows[keep], cols[keep])), shape=(mesh.num_cells,) * 2)
    for r in range(size):
        idx = np.nonzero(parts == r)[0]
        ncomp, _ = connected_components(G[idx][:, idx], directed=False)
        assert ncomp == 1
    owned = []
    for r in range(size):
        local, plan = partition_mesh(mesh, r, size, parts=parts)
        _check_local(mesh, local)
        owned += list(local.global_ids[:local.num_owned])
        # what I receive from a peer is exactly what that peer sends to me (matched by global id)
        for peer, cells in plan.recv.items():
            other, oplan = partition_mesh(mesh, peer, size, parts=parts)
            assert np.array_equal(local.global_ids[cells], other.global_ids[oplan.send[r]])
    assert sorted(owned) == list(range(mesh.num_cells))
