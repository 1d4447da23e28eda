"""Quadrilateral / hexahedral meshes and the per-cell connectivity the kernels read.

Stands in for the Firedrake mesh objects the reference is constructed with
(`UnitSquareMesh(N, N, quadrilateral=True)`, `RectangleMesh`, `ExtrudedMesh`,
`Mesh("*.msh")`; echemfem/solver.py:86, tests/*.py, examples/*.py).

Reference cell [0,1]^d; local vertices lexicographic, last coordinate fastest;
local facet 2k+s is the face xi_k = s.  Box meshes carry
Firedrake's boundary ids 1: x=0, 2: x=1, 3: y=0, 4: y=1 (5/6: z).

The kernels never see a facet list.  DG assembly is owner-computes: every cell
evaluates its own 2d facets, so connectivity is stored per cell:

  nbr[c, lf]      neighbour cell, or -(1 + marker_class) on the boundary
  nbr_code[c, lf] neighbour's local facet (3 bits) | orientation (3 bits) << 3
  slot[c, lf]     rank of the neighbour among the sorted cells {c} U nbrs(c)
                  (position of its column block inside a CSR row of c)
  slot_self[c], ncol_cells[c]
"""
import numpy as np

from . import symbolic


def _facet_local_vertices(d, lf):
    k, s = divmod(lf, 2)
    return [v for v in range(2 ** d) if ((v >> (d - 1 - k)) & 1) == s]


class Mesh:
    def __init__(self, coords, cells, marker_of_facet=None, name="mesh", layers=None):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int64)
        self.dim = self.coords.shape[1]
        self.num_cells = self.cells.shape[0]
        self.name = name
        self.layers = layers
        self._marker_of_facet = marker_of_facet
        symbolic.set_geometric_dimension(self.dim)
        self._connect()

    @property
    def coordinates(self):
        """`mesh.coordinates.dat.data[:] = ...` (examples/catalyst_layer_Cu_full.py:187-188): a writable view of
        the vertex coordinates, shape (nv,) in 1D and (nv, d) otherwise.  Vertices of tensor meshes are numbered
        lexicographically (left to right in 1D).  Move vertices before the solver is set up."""
        view = self.coords[:, 0] if self.dim == 1 else self.coords
        dim = self.dim

        class _Dat:
            data = view
            data_ro = view

        class _Coords:
            dat = _Dat

            def __getitem__(self, k):
                # `mesh.coordinates[k]` inside expressions (examples/BMCSL.py:99): the k-th spatial coordinate
                return symbolic.coordinate(dim)[k]

            def __iter__(self):
                return iter(symbolic.coordinate(dim))
        return _Coords()

    # Firedrake-compatible queries used by the reference (solver.py:202, 221, 752)
    def topological_dimension(self):
        return self.dim

    def geometric_dimension(self):
        return self.dim

    @property
    def cell_set(self):
        class _S:
            size = self.num_cells
        return _S

    # -- topology -------------------------------------------------------------
    def _connect(self):
        d = self.dim
        nlf = 2 * d
        nc = self.num_cells
        nfv = 2 ** (d - 1)
        fv = np.empty((nc, nlf, nfv), dtype=np.int64)
        for lf in range(nlf):
            fv[:, lf, :] = self.cells[:, _facet_local_vertices(d, lf)]
        key = np.sort(fv, axis=2).reshape(nc * nlf, nfv)
        # hash facets: lexsort rows, neighbours become adjacent
        order = np.lexsort(key.T[::-1])
        ks = key[order]
        same = np.all(ks[1:] == ks[:-1], axis=1)
        a = order[:-1][same]
        b = order[1:][same]
        nbr = np.full(nc * nlf, -1, dtype=np.int64)
        nbr_lf = np.zeros(nc * nlf, dtype=np.int64)
        nbr[a] = b // nlf
        nbr[b] = a // nlf
        nbr_lf[a] = b % nlf
        nbr_lf[b] = a % nlf
        self.nbr_cell = nbr.reshape(nc, nlf)
        self.nbr_lf = nbr_lf.reshape(nc, nlf)
        self._fv = fv
        # orientation of the neighbour's facet parametrisation relative to ours
        self.nbr_orient = np.zeros((nc, nlf), dtype=np.int64)
        if d >= 2:
            self._orientations()
        # exterior markers
        self.ext_marker = np.zeros((nc, nlf), dtype=np.int64)
        ext = np.argwhere(self.nbr_cell < 0)
        if len(ext):
            fcoords = self.coords[fv[ext[:, 0], ext[:, 1]]]            # (ne, nfv, d)
            if self._marker_of_facet is not None:
                self.ext_marker[ext[:, 0], ext[:, 1]] = self._marker_of_facet(fcoords, fv[ext[:, 0], ext[:, 1]])
        self.num_interior_facets = int((self.nbr_cell >= 0).sum() // 2)
        self.num_exterior_facets = int(len(ext))

    def _orientations(self):
        """3-bit code: bit0 swap of the two facet axes (3D), bit1/bit2 flip of my facet axis 0/1."""
        d = self.dim
        nlf = 2 * d
        cs, lfs = np.nonzero(self.nbr_cell >= 0)
        mine = self._fv[cs, lfs]                                        # (nf, nfv) my facet vertices
        theirs = self._fv[self.nbr_cell[cs, lfs], self.nbr_lf[cs, lfs]]
        # position, in their facet-local numbering, of each of my facet vertices
        pos = np.argmax(mine[:, :, None] == theirs[:, None, :], axis=2)  # (nf, nfv)
        if d == 2:
            code = (pos[:, 0] == 1).astype(np.int64) << 1               # reversed edge
        else:
            # facet-local vertex m = 2*t0 + t1 ; my axis 0 moves m by 2, axis 1 by 1
            p0, p1, p2 = pos[:, 0], pos[:, 1], pos[:, 2]
            d0 = p2 ^ p0                                                # their bit pattern toggled by my axis 0
            swap = (d0 == 1).astype(np.int64)                           # my axis 0 -> their axis 1
            # their coordinates of my origin vertex: (p0 >> 1, p0 & 1)
            t0o, t1o = p0 >> 1, p0 & 1
            flip0 = np.where(swap == 1, t1o, t0o)                       # flip along the axis my axis 0 maps to
            flip1 = np.where(swap == 1, t0o, t1o)
            code = swap | (flip0 << 1) | (flip1 << 2)
        self.nbr_orient[cs, lfs] = code

    # -- arrays handed to the device ------------------------------------------
    def cell_coords(self):
        return np.ascontiguousarray(self.coords[self.cells])            # (nc, 2^d, d)

    def connectivity(self, marker_class):
        """(nbr int32, code uint8, slot uint8, slot_self uint8, ncol uint8)."""
        nc, nlf = self.nbr_cell.shape
        nbr = self.nbr_cell.copy()
        lut = np.zeros(int(self.ext_marker.max()) + 2 if nc else 1, dtype=np.int64)
        for m, k in marker_class.items():
            if 0 <= int(m) < len(lut):
                lut[int(m)] = k
        cls = lut[self.ext_marker]
        bnd = nbr < 0
        nbr[bnd] = -(1 + cls[bnd])
        code = (self.nbr_lf | (self.nbr_orient << 3)).astype(np.uint8)
        big = np.iinfo(np.int64).max
        allc = np.concatenate([np.arange(nc)[:, None], np.where(bnd, big, self.nbr_cell)], axis=1)
        rank = np.argsort(np.argsort(allc, axis=1, kind="stable"), axis=1, kind="stable")
        slot_self = rank[:, 0].astype(np.uint8)
        slot = rank[:, 1:].astype(np.uint8)
        ncol = (1 + (~bnd).sum(axis=1)).astype(np.uint8)
        return nbr.astype(np.int32), code, slot, slot_self, ncol


def _box_markers(lo, hi):
    lo = np.asarray(lo, dtype=float)
    hi = np.asarray(hi, dtype=float)
    tol = 1e-12 * max(1.0, float(np.max(np.abs(hi - lo))))

    def fn(fcoords, fverts):
        out = np.zeros(fcoords.shape[0], dtype=np.int64)
        for k in range(len(lo)):
            on_lo = np.all(np.abs(fcoords[:, :, k] - lo[k]) < tol, axis=1)
            on_hi = np.all(np.abs(fcoords[:, :, k] - hi[k]) < tol, axis=1)
            out[on_lo & (out == 0)] = 2 * k + 1
            out[on_hi & (out == 0)] = 2 * k + 2
        return out
    return fn


def tile_order(shape, tile):
    """Permutation of the lexicographic cell numbers of a tensor mesh that makes the cells of one tile
    (a box of `tile` cells per direction) contiguous: tiles in lexicographic order, cells inside a tile
    lexicographic.  Contiguous cell ranges are what the block-Jacobi/ILU(0) preconditioner and the
    one-warp-per-batch assembly kernels work on, so compact tiles give compact blocks (the reference's
    DMPlex numbering is reordered for locality as well; SURVEY.md 8c Q7 -- numbering is not part of parity)."""
    shape = tuple(int(n) for n in shape)
    d = len(shape)
    tile = tuple(int(t) for t in (tile if np.ndim(tile) else (tile,) * d))
    idx = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    keys = [i.reshape(-1) % t for i, t in zip(idx, tile)][::-1] + [i.reshape(-1) // t for i, t in zip(idx, tile)][::-1]
    return np.lexsort(keys)          # last key (tile index of direction 0) is the primary one


def TensorMesh(axes, name="tensor", marker_of_facet=None, layers=None, tile=None):
    """Tensor-product mesh from one vertex-coordinate array per direction (graded meshes).

    tile: cells are numbered tile by tile (see `tile_order`) instead of lexicographically."""
    axes = [np.asarray(a, dtype=float) for a in axes]
    d = len(axes)
    nv = [len(a) for a in axes]
    grid = np.meshgrid(*axes, indexing="ij")
    coords = np.stack([g.reshape(-1) for g in grid], axis=1)
    vid = np.arange(int(np.prod(nv))).reshape(nv)
    corner = []
    for v in range(2 ** d):
        sl = tuple(slice(((v >> (d - 1 - m)) & 1), nv[m] - 1 + ((v >> (d - 1 - m)) & 1)) for m in range(d))
        corner.append(vid[sl].reshape(-1))
    cells = np.stack(corner, axis=1)
    if marker_of_facet is None:
        marker_of_facet = _box_markers([a[0] for a in axes], [a[-1] for a in axes])
    shape = tuple(n - 1 for n in nv)
    perm = None
    if tile is not None:
        perm = tile_order(shape, tile)
        cells = cells[perm]
    m = Mesh(coords, cells, marker_of_facet, name=name, layers=layers)
    m.shape = shape
    m.axes = axes
    m.cell_perm = perm               # new cell i is lexicographic cell perm[i]
    if tile is not None:
        t = tuple(int(v) for v in (tile if np.ndim(tile) else (tile,) * d))
        if all(n % k == 0 for n, k in zip(shape, t)):
            m.block_hint = int(np.prod(t))      # cells per preconditioner block = one tile
            m.tile_shape = t
    return m


def UnitSquareMesh(nx, ny, quadrilateral=True, tile=None, **kw):
    if not quadrilateral:
        raise NotImplementedError("only quadrilateral/hexahedral cells are supported")
    return TensorMesh([np.linspace(0.0, 1.0, nx + 1), np.linspace(0.0, 1.0, ny + 1)], name="unit_square", tile=tile)


def RectangleMesh(nx, ny, Lx, Ly, quadrilateral=True, **kw):
    if not quadrilateral:
        raise NotImplementedError("only quadrilateral/hexahedral cells are supported")
    return TensorMesh([np.linspace(0.0, Lx, nx + 1), np.linspace(0.0, Ly, ny + 1)], name="rectangle")


def UnitCubeMesh(nx, ny, nz, hexahedral=True, **kw):
    if not hexahedral:
        raise NotImplementedError("only quadrilateral/hexahedral cells are supported")
    return TensorMesh([np.linspace(0.0, 1.0, n + 1) for n in (nx, ny, nz)], name="unit_cube")


def ExtrudedMesh(base, layers, layer_height=None, **kw):
    """Extrusion of a tensor-product quad mesh into hexes (examples/mms_scaling.py:72-76).

    Side boundaries keep the base ids (the reference integrates `ds_v` only,
    solver.py:221-228); top and bottom get ids 5/6 which no form references.
    """
    if layer_height is None:
        layer_height = 1.0 / layers
    if not hasattr(base, "axes"):
        xs = np.unique(base.coords[:, 0])
        ys = np.unique(base.coords[:, 1])
        if len(xs) * len(ys) != base.coords.shape[0]:
            raise NotImplementedError("ExtrudedMesh needs a tensor-product base mesh")
        axes = [xs, ys]
    else:
        axes = base.axes
    z = np.arange(layers + 1) * float(layer_height)
    return TensorMesh(list(axes) + [z], name=base.name + "_extruded", layers=layers,
                      marker_of_facet=_extruded_markers(base, float(z[0]), float(z[-1])))


def _extruded_markers(base, z0, z1):
    """Side facets of the extrusion carry the id of the base-mesh edge below them; bottom / top get 5 / 6."""
    base_fn = base._marker_of_facet
    if base_fn is None or base.dim != 2:
        return None                                     # TensorMesh falls back to box ids
    tol = 1e-12 * max(1.0, abs(z1 - z0))

    def fn(fcoords, fverts):
        out = np.zeros(fcoords.shape[0], dtype=np.int64)
        zc = fcoords[:, :, 2]
        flat = np.abs(zc - zc[:, :1]).max(axis=1) < tol
        out[flat & (np.abs(zc[:, 0] - z0) < tol)] = 5
        out[flat & (np.abs(zc[:, 0] - z1) < tol)] = 6
        side = ~flat
        if side.any():
            # local facet vertices are lexicographic with z fastest: vertices 0 and 2 are the two base points
            edge = fcoords[side][:, [0, 2], :2]
            out[side] = base_fn(edge, None)
        return out
    return fn


def _shape_derivs(d, xi):
    """Vertex shape functions and reference derivatives at points xi (npts, d)."""
    npts = xi.shape[0]
    N = np.ones((npts, 2 ** d))
    dN = np.ones((npts, 2 ** d, d))
    for v in range(2 ** d):
        for k in range(d):
            bit = (v >> (d - 1 - k)) & 1
            val = xi[:, k] if bit else 1.0 - xi[:, k]
            N[:, v] *= val
            for l in range(d):
                dN[:, v, l] *= (1.0 if bit else -1.0) if l == k else val
    return N, dN


def cell_geometry(mesh, chunk=1 << 18):
    """CellVolume, FacetArea (per cell and local facet) and CellDiameter.

    Volumes and areas are integrated with a 3-point Gauss rule per direction,
    exact for multilinear cells with planar facets; the diameter is the largest
    vertex distance.
    """
    from .tables import gauss_points
    d = mesh.dim
    gp, gw = gauss_points(3)
    grids = np.meshgrid(*([np.arange(3)] * d), indexing="ij")
    qi = np.stack([g.reshape(-1) for g in grids], axis=1)
    xi = gp[qi]
    w = np.prod(gw[qi], axis=1)
    _, dN = _shape_derivs(d, xi)
    X = mesh.cell_coords()
    nc = mesh.num_cells
    vol = np.empty(nc)
    area = np.empty((nc, 2 * d))
    diam = np.empty(nc)
    if d > 1:
        gridf = np.meshgrid(*([np.arange(3)] * (d - 1)), indexing="ij")
        qf = np.stack([g.reshape(-1) for g in gridf], axis=1)
        wf = np.prod(gw[qf], axis=1)
    else:
        qf = np.zeros((1, 0), dtype=int)
        wf = np.ones(1)
    for s0 in range(0, nc, chunk):
        Xc = X[s0:s0 + chunk]
        J = np.einsum("qvl,cvk->cqkl", dN, Xc)
        vol[s0:s0 + chunk] = np.einsum("q,cq->c", w, np.abs(np.linalg.det(J)))
        for lf in range(2 * d):
            k, s = divmod(lf, 2)
            xif = np.empty((qf.shape[0], d))
            xif[:, k] = float(s)
            rest = [m for m in range(d) if m != k]
            for a, m in enumerate(rest):
                xif[:, m] = gp[qf[:, a]]
            _, dNf = _shape_derivs(d, xif)
            Jf = np.einsum("qvl,cvk->cqkl", dNf, Xc)
            Jinv = np.linalg.inv(Jf)
            nrm = np.sqrt((Jinv[..., k, :] ** 2).sum(-1))
            area[s0:s0 + chunk, lf] = np.einsum("q,cq->c", wf, np.abs(np.linalg.det(Jf)) * nrm)
        diff = Xc[:, :, None, :] - Xc[:, None, :, :]
        diam[s0:s0 + chunk] = np.sqrt((diff ** 2).sum(-1)).max(axis=(1, 2))
    return vol, area, diam
