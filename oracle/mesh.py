"""Quadrilateral / hexahedral meshes for the oracle.

Conventions: reference cell [0,1]^d, vertices numbered
lexicographically with the last coordinate fastest ((0,0),(0,1),(1,0),(1,1));
local facet 2*k+s is the face x_k = s.  Boundary ids of the generated unit
square / box follow Firedrake's `UnitSquareMesh`: 1: x=0, 2: x=1, 3: y=0,
4: y=1 (5: z=0, 6: z=1).  Cell numbering is lexicographic (the reference's
DMPlex numbering is not reproducible offline, Q7).
"""
import numpy as np


class Mesh:
    """Conforming mesh of bilinear quads / trilinear hexes."""

    def __init__(self, coords, cells, boundary_id_fn=None, facet_markers=None):
        self.coords = np.asarray(coords, dtype=float)
        self.cells = np.asarray(cells, dtype=np.int64)
        self.dim = self.coords.shape[1]
        self.num_cells = self.cells.shape[0]
        assert self.cells.shape[1] == 2 ** self.dim
        self._build_facets(boundary_id_fn, facet_markers)

    # -- topology ---------------------------------------------------------
    def local_facet_vertices(self, lf):
        """Local vertex numbers of local facet lf, in facet-lexicographic order."""
        d = self.dim
        k, s = divmod(lf, 2)
        out = []
        for v in range(2 ** d):
            bits = [(v >> (d - 1 - m)) & 1 for m in range(d)]
            if bits[k] == s:
                out.append(v)
        return out

    def _build_facets(self, boundary_id_fn, facet_markers):
        d = self.dim
        nlf = 2 * d
        nc = self.num_cells
        keys = np.empty((nc, nlf, 2 ** (d - 1)), dtype=np.int64)
        for lf in range(nlf):
            lv = self.local_facet_vertices(lf)
            keys[:, lf, :] = np.sort(self.cells[:, lv], axis=1)
        flat = keys.reshape(nc * nlf, -1)
        _, inv, cnt = np.unique(flat, axis=0, return_inverse=True, return_counts=True)
        inv = inv.ravel()
        if cnt.max() > 2:
            raise ValueError("non-manifold facet")
        order = np.argsort(inv, kind="stable")     # (cell, lf) pairs grouped by facet
        grp = inv[order]
        is_pair_first = np.zeros(len(order), dtype=bool)
        is_pair_first[:-1] = grp[:-1] == grp[1:]
        first = order[is_pair_first]
        second = order[np.roll(is_pair_first, 1)]
        single = order[cnt[grp] == 1]
        int_f = np.stack([first // nlf, second // nlf, first % nlf, second % nlf], axis=1)
        int_f = int_f[np.lexsort((int_f[:, 2], int_f[:, 0]))]
        ext_f = np.stack([single // nlf, single % nlf], axis=1)
        ext_f = ext_f[np.lexsort((ext_f[:, 1], ext_f[:, 0]))]
        self.int_facets = int_f.astype(np.int64).reshape(-1, 4)
        self.ext_facets = ext_f.astype(np.int64).reshape(-1, 2)
        # neighbour table: nbr[c, lf] = cell or -1 ; nbr_lf[c, lf]
        self.nbr = -np.ones((self.num_cells, nlf), dtype=np.int64)
        self.nbr_lf = -np.ones((self.num_cells, nlf), dtype=np.int64)
        c0, c1, l0, l1 = self.int_facets.T
        self.nbr[c0, l0] = c1
        self.nbr_lf[c0, l0] = l1
        self.nbr[c1, l1] = c0
        self.nbr_lf[c1, l1] = l0
        # exterior markers
        self.ext_marker = np.zeros(len(self.ext_facets), dtype=np.int64)
        for f, (c, lf) in enumerate(self.ext_facets):
            lv = self.local_facet_vertices(lf)
            gv = self.cells[c, lv]
            if facet_markers is not None:
                self.ext_marker[f] = facet_markers[tuple(sorted(gv))]
            elif boundary_id_fn is not None:
                self.ext_marker[f] = boundary_id_fn(self.coords[gv])

    # -- geometry ---------------------------------------------------------
    def cell_vertex_coords(self):
        """(nc, 2^d, d) vertex coordinates per cell."""
        return self.coords[self.cells]

    def neighbour_facet_map(self, c0, l0, c1, l1):
        """Facet-local coordinates in c1 of the facet-vertices of (c0,l0).

        Returns B (2^(d-1), d-1): row m is the facet-local coordinate, in
        cell c1's facet l1, of the m-th facet vertex of (c0, l0).  The map
        t -> t' is then sum_m N_m(t) B[m] with N_m multilinear.
        """
        d = self.dim
        lv0 = self.local_facet_vertices(l0)
        lv1 = self.local_facet_vertices(l1)
        g0 = self.cells[c0, lv0]
        g1 = list(self.cells[c1, lv1])
        k1 = l1 // 2
        B = np.zeros((len(lv0), d - 1))
        for m, g in enumerate(g0):
            v1 = lv1[g1.index(g)]
            bits = [(v1 >> (d - 1 - mm)) & 1 for mm in range(d)]
            del bits[k1]
            B[m] = bits
        return B


def _box_marker(lo, hi, tol=1e-12):
    d = len(lo)

    def fn(xv):
        for k in range(d):
            if np.all(np.abs(xv[:, k] - lo[k]) < tol):
                return 2 * k + 1
            if np.all(np.abs(xv[:, k] - hi[k]) < tol):
                return 2 * k + 2
        raise ValueError("exterior facet not on the box boundary")
    return fn


def tensor_mesh(axes, boundary_id_fn=None):
    """Tensor-product mesh from 1D vertex coordinate arrays, one per direction."""
    axes = [np.asarray(a, dtype=float) for a in axes]
    d = len(axes)
    grids = np.meshgrid(*axes, indexing="ij")
    coords = np.stack([g.ravel() for g in grids], axis=1)
    nv = [len(a) for a in axes]
    ncell = [n - 1 for n in nv]
    cidx = np.stack([g.ravel() for g in np.meshgrid(*[np.arange(n) for n in ncell], indexing="ij")], axis=1)
    cells = np.zeros((cidx.shape[0], 2 ** d), dtype=np.int64)
    for v in range(2 ** d):
        bits = [(v >> (d - 1 - m)) & 1 for m in range(d)]
        vid = np.zeros(cidx.shape[0], dtype=np.int64)
        for m in range(d):
            vid = vid * nv[m] + (cidx[:, m] + bits[m])
        cells[:, v] = vid
    if boundary_id_fn is None:
        boundary_id_fn = _box_marker([a[0] for a in axes], [a[-1] for a in axes])
    return Mesh(coords, cells, boundary_id_fn=boundary_id_fn)


def unit_square_mesh(nx, ny=None):
    """`UnitSquareMesh(nx, ny, quadrilateral=True)` (tests/test_*.py)."""
    ny = nx if ny is None else ny
    return tensor_mesh([np.linspace(0, 1, nx + 1), np.linspace(0, 1, ny + 1)])


def unit_cube_mesh(nx, ny, nz):
    return tensor_mesh([np.linspace(0, 1, nx + 1), np.linspace(0, 1, ny + 1),
                        np.linspace(0, 1, nz + 1)])
