"""Host-side pieces of the continuous-Q1 auxiliary space (echemfem_b200/auxspace.py): the (cell, corner) -> vertex
table and the embedding block against the node coordinates of the DG space, and the coordinate-based linear
interpolation between vertex sets."""
import numpy as np
import pytest

from echemfem_b200.auxspace import dg_to_cg_nodes, embedding_block, interpolation_1d
from echemfem_b200.function import FunctionSpace
from echemfem_b200.mesh import TensorMesh, ExtrudedMesh
from echemfem_b200.multigrid import Hierarchy


@pytest.mark.parametrize("tile", [None, 4, (8, 2)])
def test_dg_dofs_map_to_their_vertices(tile):
    ax = [np.linspace(0.0, 2.0, 17) ** 1.3, np.linspace(-1.0, 1.0, 13)]
    mesh = TensorMesh(ax, tile=tile)
    V = FunctionSpace(mesh, "DG", 1)
    h = Hierarchy(mesh, 1, tile=tile if tile is not None else 4)
    nrows = V.dim() + 8                                    # ghost rows behind the owned ones stay unmapped
    vtab = dg_to_cg_nodes(mesh.shape, h.levels[0].inv, nrows)
    assert np.all(vtab[V.dim():] == -1) and np.all(vtab[:V.dim()] >= 0)
    g = np.meshgrid(*ax, indexing="ij")
    verts = np.stack([a.reshape(-1) for a in g], axis=1)
    # the embedding of the coordinate functions (continuous, linear per direction) gives the DG node coordinates
    W = embedding_block(2)
    np.testing.assert_allclose(W.sum(axis=1), 1.0, rtol=0, atol=1e-15)
    corners = verts[vtab[:V.dim()].reshape(-1, 4)]               # (cells, corner, 2)
    np.testing.assert_allclose(np.einsum("av,cvk->cak", W, corners).reshape(-1, 2), V.node_coords(), rtol=0, atol=1e-14)
    # every vertex is a corner of as many cells as surround it
    hits = np.bincount(vtab[:V.dim()], minlength=len(verts)).reshape(len(ax[0]), len(ax[1]))
    assert hits[0, 0] == 1 and hits[0, 1] == 2 and hits[1, 1] == 4


def test_dg_dofs_map_to_their_vertices_3d():
    base = TensorMesh([np.linspace(0, 1, 5), np.linspace(0, 2, 7)])
    mesh = ExtrudedMesh(base, 4, 0.25)
    if not hasattr(mesh, "axes"):
        pytest.skip("extruded mesh does not expose tensor axes")
    V = FunctionSpace(mesh, "DG", 1)
    h = Hierarchy(mesh, 1, tile=2)
    vtab = dg_to_cg_nodes(mesh.shape, h.levels[0].inv, V.dim())
    g = np.meshgrid(*mesh.axes, indexing="ij")
    verts = np.stack([a.reshape(-1) for a in g], axis=1)
    corners = verts[vtab.reshape(-1, 8)]
    np.testing.assert_allclose(np.einsum("av,cvk->cak", embedding_block(3), corners).reshape(-1, 3), V.node_coords(),
                               rtol=0, atol=1e-14)


@pytest.mark.parametrize("n", [8, 9, 11])
def test_interpolation_reproduces_linear_functions(n):
    ax_f = np.linspace(0.0, 1.0, n + 1) ** 1.5
    ax_c = ax_f[::2] if n % 2 == 0 else np.append(ax_f[::2], ax_f[-1])
    P = interpolation_1d(ax_f, ax_c)
    assert P.shape == (n + 1, len(ax_c))
    np.testing.assert_allclose(P @ (3.0 * ax_c - 1.0), 3.0 * ax_f - 1.0, rtol=0, atol=1e-14)
    np.testing.assert_allclose(np.asarray(P.sum(axis=1)).ravel(), 1.0, rtol=0, atol=1e-14)
    assert P.nnz <= 2 * (n + 1) and P.data.min() > 0.0


def test_coarse_embedding_on_two_strips_agrees_on_shared_vertices():
    """The coarse space shared by all ranks is addressed by coordinates: two strips of one mesh (ghost layers
    included) embed the same global vertex function, each reproducing it at its own DG nodes."""
    from echemfem_b200.auxspace import coarse_embedding
    from echemfem_b200.partition import strip_partition
    axes = [np.linspace(0.0, 2.0, 33) ** 1.2, np.linspace(0.0, 1.0, 17)]
    cax = [a[::4] for a in axes]
    g = np.meshgrid(*cax, indexing="ij")
    f = (1.5 * g[0] - 0.5 * g[1] + g[0] * g[1]).ravel()            # bilinear per coarse cell: in the coarse space
    seen = np.zeros(f.size, dtype=int)
    for rank in range(2):
        mesh, _ = strip_partition(axes, rank, 2, tile=8)
        assert mesh.global_axes is axes or all(np.array_equal(a, b) for a, b in zip(mesh.global_axes, axes))
        V = FunctionSpace(mesh, "DG", 1)
        x = V.node_coords()
        cI, W, vert = coarse_embedding(x, 4, cax)
        assert cI.shape == (mesh.num_cells, 2) and W.shape == (len(x), 4)
        np.testing.assert_allclose(W.sum(axis=1), 1.0, atol=1e-14)
        # the coarse function is bilinear on each coarse cell only in the reference coordinates of the coarse cell;
        # with a graded axis check the two affine parts and the product through the coarse cell's own coordinates
        np.testing.assert_allclose((W * g[0].ravel()[vert]).sum(axis=1), x[:, 0], atol=1e-13)
        np.testing.assert_allclose((W * g[1].ravel()[vert]).sum(axis=1), x[:, 1], atol=1e-13)
        np.testing.assert_allclose((W * (g[0] * g[1]).ravel()[vert]).sum(axis=1), x[:, 0] * x[:, 1], atol=1e-13)
        seen[np.unique(vert)] += 1
    # vertices on and next to the cut are used by both ranks (owned cells of one, ghost cells of the other)
    assert seen.min() >= 1 and (seen == 2).sum() >= 2 * len(cax[1])


def test_hierarchy_with_elongated_tiles():
    """Per-direction tiles (blocks elongated along the flow): the coarse levels keep the tile shape while it
    divides them and halve it per direction where it does not; prolongations stay exact embeddings."""
    from echemfem_b200.multigrid import _fit_tile
    assert _fit_tile((64, 64), 8) == 8 and _fit_tile((12, 64), 8) is None          # scalar tiles: all or nothing
    assert _fit_tile((64, 64), (32, 2)) == (32, 2)
    assert _fit_tile((16, 8), (32, 2)) == (16, 2)
    assert _fit_tile((3, 3), (32, 2)) is None
    ax = [np.linspace(0.0, 1.0, 65), np.linspace(0.0, 1.0, 33)]
    mesh = TensorMesh(ax, tile=(16, 4))
    assert mesh.tile_shape == (16, 4) and mesh.block_hint == 64
    h = Hierarchy(mesh, 1, tile=(16, 4), coarsest_cells=16)
    assert [L.shape for L in h.levels][:3] == [(64, 32), (32, 16), (16, 8)]
    assert h.levels[0].ctile == (16, 4) and h.levels[1].ctile == (16, 4)
    V = FunctionSpace(mesh, "DG", 1)
    x = V.node_coords()
    P = h.P[0]
    # coarse level 1 numbered in its own tiles: compare through a function both levels represent exactly
    cm = TensorMesh(h.levels[1].axes, tile=h.levels[0].ctile)
    xc = FunctionSpace(cm, "DG", 1).node_coords()
    f = lambda p: 1.0 + 2.0 * p[:, 0] - p[:, 1] + 0.5 * p[:, 0] * p[:, 1]
    np.testing.assert_allclose(P @ f(xc), f(x), rtol=0, atol=1e-13)
