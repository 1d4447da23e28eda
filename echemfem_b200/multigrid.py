"""Geometric multigrid preconditioner for DG systems on tensor-product meshes (the reference's `gmg` option set).

Reference: `init_solver_parameters("gmg")` (echemfem/solver.py:1190-1201): FGMRES around a multigrid V-cycle with
one Richardson iteration of block-Jacobi/ILU per level and a direct coarse solve, on a Firedrake `MeshHierarchy`
(examples/mms_scaling.py:72-76).  Here:

* the hierarchy comes from the tensor structure of the mesh: every other vertex is dropped per direction and
  level (the inverse of `MeshHierarchy`'s uniform refinement), also for graded axes and odd cell counts (the
  297 x 49 mesh of examples/bortels_threeion_nondim.py);
* prolongation is the natural embedding of the coarse DG space into the fine one (the coarse polynomial of the
  parent cell evaluated at the fine nodes), a tensor product of 1D matrices assembled on the host; restriction
  is its transpose;
* coarse operators are Galerkin products `P^T A P` of the assembled Jacobian (CSR x CSR products through
  cuSPARSE/torch: a library call in the SETUP, once per Newton iteration) -- no re-assembly, so user physics and
  boundary conditions need no coarse counterparts;
* the smoother is the same block-Jacobi/ILU(0) as on the fine level (`ech_bjacobi_ilu0_*`, blocks = contiguous
  ranges of tile-numbered cells), the transfer operators are applied with `ech_spmv`, the coarsest system is
  solved with a dense inverse.

Everything per level lives in HBM; a V-cycle is a fixed sequence of kernel launches.
"""
import os

import numpy as np
import scipy.sparse as sp
import torch

from . import capi
from .capi import ptr
from .mesh import tile_order
from .tables import element_nodes, basis_table


def prolongation_1d(ax_f, ax_c, p):
    """((n_f * n1) x (n_c * n1)) matrix of one direction: fine dof (cell i, node a) <- coarse dof (cell i // 2, node B)."""
    nodes = element_nodes(p, "DG")
    n1 = p + 1
    nf, nc = len(ax_f) - 1, len(ax_c) - 1
    i = np.arange(nf)
    I = i // 2
    x = ax_f[:-1][:, None] + nodes[None, :] * np.diff(ax_f)[:, None]               # (nf, n1) fine node positions
    xi = (x - ax_c[I][:, None]) / (ax_c[I + 1] - ax_c[I])[:, None]
    L, _ = basis_table(nodes, xi.reshape(-1))                                        # (nf * n1, n1)
    rows = np.repeat(np.arange(nf * n1), n1)
    cols = (np.repeat(I, n1)[:, None] * n1 + np.arange(n1)[None, :]).reshape(-1)
    return sp.csr_matrix((L.reshape(-1), (rows, cols)), shape=(nf * n1, nc * n1))


def _dof_map(shape, n1, inv_perm):
    """kron index (d_0, d_1, ...) with d_k = i_k * n1 + a_k  ->  dof (cell number, lexicographic node)."""
    d = len(shape)
    sizes = [s * n1 for s in shape]
    idx = np.arange(int(np.prod(sizes)))
    cell = np.zeros_like(idx)
    node = np.zeros_like(idx)
    rem = idx
    parts = []
    for k in range(d - 1, -1, -1):
        parts.append(rem % sizes[k])
        rem = rem // sizes[k]
    parts = parts[::-1]
    for k in range(d):
        cell = cell * shape[k] + parts[k] // n1
        node = node * n1 + parts[k] % n1
    if inv_perm is not None:
        cell = inv_perm[cell]
    return cell * (n1 ** d) + node


class Level:
    pass


def dg_pattern(nf, n, ncell, cells_sorted, ncol, own, nbrm, nrows_per_field=None):
    """Field-blocked AIJ pattern of a DG space: row (i, c, a) holds, for every coupled field j,
    one n-entry block per cell of sorted({c} U neighbours(c)) if the coupling crosses facets, else the own block.
    cells_sorted: (ncell, 2d+1) cell ids in ascending order, -1 padded; returns (rowptr int64, colidx int32)."""
    N = nrows_per_field or ncell * n
    ncol = np.asarray(ncol, dtype=np.int64)
    self_pos = np.argmax(cells_sorted == np.arange(ncell)[:, None], axis=1)
    rowptr_parts, col_parts = [], []
    base = 0
    for i in range(nf):
        nblk = np.zeros(ncell, dtype=np.int64)
        for j in range(nf):
            if own[i, j]:
                nblk += ncol if nbrm[i, j] else 1
        # block starts per cell, in row order
        tot = int(nblk.sum())
        starts = np.empty(tot, dtype=np.int64)
        cell_off = np.concatenate([[0], np.cumsum(nblk)])[:-1]
        seg = np.zeros(ncell, dtype=np.int64)
        for j in range(nf):
            if not own[i, j]:
                continue
            if nbrm[i, j]:
                for p in range(cells_sorted.shape[1]):
                    m = p < ncol
                    starts[(cell_off + seg + p)[m]] = j * N + cells_sorted[m, p] * n
                seg += ncol
            else:
                starts[cell_off + seg] = j * N + np.arange(ncell) * n
                seg += 1
        # expand: every row a of the cell repeats the cell's block list; every block is n consecutive columns
        lens_rows = np.repeat(nblk * n, n)                                        # (ncell * n,)
        rowptr_parts.append(base + np.concatenate([[0], np.cumsum(lens_rows)])[:-1])
        # (cell, a, block) order: position of block q of cell c in copy a = n * cell_off[c] + a * nblk[c] + q
        q_in_cell = np.arange(tot) - np.repeat(cell_off, nblk)
        out_base = np.repeat(np.concatenate([[0], np.cumsum(nblk * n)])[:-1], nblk)   # in blocks, per cell, all copies
        blocks_all = np.empty(tot * n, dtype=np.int64)
        for a in range(n):
            blocks_all[out_base + a * np.repeat(nblk, nblk) + q_in_cell] = starts
        cols = (blocks_all[:, None] + np.arange(n)[None, :]).reshape(-1)
        col_parts.append(cols.astype(np.int32))
        base += int(lens_rows.sum())
    rowptr = np.concatenate(rowptr_parts + [[base]]).astype(np.int64)
    return rowptr, np.concatenate(col_parts)


def _sorted_cells(nbr):
    """(ncell, 2d+1) ascending cell ids of {c} U valid neighbours, -1 padded, and the count."""
    nbr = np.asarray(nbr, dtype=np.int64)
    nc = nbr.shape[0]
    big = np.iinfo(np.int64).max
    allc = np.concatenate([np.arange(nc, dtype=np.int64)[:, None], np.where(nbr >= 0, nbr, big)], axis=1)
    allc.sort(axis=1)
    ncol = (allc < big).sum(axis=1)
    allc[allc == big] = -1
    return allc, ncol


def _fit_tile(shape, tile):
    """Tile of a coarse level: `tile` itself if it divides the level (None -> lexicographic numbering otherwise);
    a per-direction tile (tuple, e.g. elongated along the flow) is halved per direction until it divides."""
    if not np.ndim(tile):
        return tile if all(s % tile == 0 and s > tile for s in shape) else None
    t = [int(v) for v in tile]
    for k, s in enumerate(shape):
        while t[k] > 1 and (s % t[k] or s < t[k]):
            t[k] //= 2
    return tuple(t) if int(np.prod(t)) > 1 else None


class Hierarchy:
    """Levels of a tensor mesh and the scalar prolongations between them (host side)."""

    def __init__(self, mesh, degree, tile=16, coarsest_cells=64, max_levels=12, fine_rows=None):
        """fine_rows: rows of the finest prolongation (>= fine dofs): a rank's vectors also hold its ghost cells
        (after the owned ones), which the hierarchy of the owned cells leaves untouched."""
        if hasattr(mesh, "owned_axes"):
            mesh_axes, perm = mesh.owned_axes, getattr(mesh, "cell_perm", None)   # one rank's strip (strip_partition)
        elif hasattr(mesh, "axes") and getattr(mesh, "halo_plan", None) is None:
            mesh_axes, perm = mesh.axes, getattr(mesh, "cell_perm", None)
        else:
            raise NotImplementedError("multigrid needs a tensor-product mesh (whole, or one strip per rank)")
        self.p = degree
        n1 = degree + 1
        d = len(mesh_axes)
        self.n = n1 ** d
        axes = [np.asarray(a, dtype=float) for a in mesh_axes]
        self.levels = []
        self.P = []                      # P[l]: level l (fine)  <-  level l + 1 (coarse), scalar dofs
        while True:
            shape = tuple(len(a) - 1 for a in axes)
            L = Level()
            L.axes, L.shape, L.perm = axes, shape, perm
            L.ncell = int(np.prod(shape))
            L.inv = None
            if perm is not None:
                L.inv = np.empty_like(perm)
                L.inv[perm] = np.arange(len(perm))
            self.levels.append(L)
            if len(self.levels) >= max_levels or L.ncell <= coarsest_cells or any(s < 4 for s in shape):
                break
            if any(s % 2 for s in shape) and L.ncell <= 256:
                break           # an odd level small enough for the dense coarse solve is the coarsest one
            # every other vertex is dropped; with an odd number of cells the last one keeps itself as parent
            # (prolongation_1d works from the coordinates, so its block is the identity)
            cax = [a[::2] if (len(a) - 1) % 2 == 0 else np.append(a[::2], a[-1]) for a in axes]
            cshape = tuple(len(a) - 1 for a in cax)
            ctile = _fit_tile(cshape, tile)
            cperm = tile_order(cshape, ctile) if ctile is not None else None
            cinv = None
            if cperm is not None:
                cinv = np.empty_like(cperm)
                cinv[cperm] = np.arange(len(cperm))
            P = prolongation_1d(axes[0], cax[0], degree)
            for k in range(1, d):
                P = sp.kron(P, prolongation_1d(axes[k], cax[k], degree), format="csr")
            P = P.tocoo()
            rf = _dof_map(shape, n1, L.inv)
            rc = _dof_map(cshape, n1, cinv)
            # parent (coarse cell, coarse numbering) of every fine cell (fine numbering)
            lexf = np.arange(L.ncell) if perm is None else np.asarray(perm)
            multi = np.unravel_index(lexf, shape)
            lexc = np.ravel_multi_index(tuple(m // 2 for m in multi), cshape)
            L.parent = (lexc if cinv is None else cinv[lexc]).astype(np.int32)
            L.cshape, L.ctile = cshape, ctile
            nrow = L.ncell * self.n
            if len(self.levels) == 1 and fine_rows is not None:
                nrow = max(nrow, int(fine_rows))
            self.P.append(sp.csr_matrix((P.data, (rf[P.row], rc[P.col])),
                                        shape=(nrow, int(np.prod(cshape)) * self.n)))
            axes, perm = cax, cperm

    @property
    def nlevels(self):
        return len(self.levels)


def _dev_csr(M, dev):
    M = M.tocsr()
    M.sort_indices()
    return (torch.from_numpy(M.indptr.astype(np.int64)).to(dev), torch.from_numpy(M.indices.astype(np.int32)).to(dev),
            torch.from_numpy(M.data.astype(np.float64)).to(dev))


def spgemm_rows(rowptr, colidx, vals, ncols, B, bnnz_per_row, max_products=1.5e8):
    """C = A B for a CSR matrix A given by (rowptr int64, colidx int32, vals) and a torch CSR tensor B, computed in
    row chunks of A (cuSPARSE SpGEMM runs out of work space on products with billions of intermediate terms).
    Returns (rowptr int64, colidx int32, vals) of C with sorted columns."""
    nrows = rowptr.numel() - 1
    nnz = int(vals.numel())
    idt = B.crow_indices().dtype
    per_row = max(1.0, nnz / max(nrows, 1)) * max(1.0, bnnz_per_row)
    chunk = int(max(1024, min(nrows, max_products / per_row)))
    try:
        return _spgemm_chunks(rowptr, colidx, vals, ncols, B, idt, nrows, chunk)
    except RuntimeError as err:                       # cuSPARSE work space: retry with smaller row chunks
        if "insufficient resources" not in str(err) and "out of memory" not in str(err):
            raise
        torch.cuda.empty_cache()
        return _spgemm_chunks(rowptr, colidx, vals, ncols, B, idt, nrows, max(256, chunk // 8))


def _spgemm_chunks(rowptr, colidx, vals, ncols, B, idt, nrows, chunk):
    crows, cols, vs = [], [], []
    base = 0
    for r0 in range(0, nrows, chunk):
        r1 = min(nrows, r0 + chunk)
        p0, p1 = int(rowptr[r0].item()), int(rowptr[r1].item())
        A = torch.sparse_csr_tensor((rowptr[r0:r1 + 1] - p0).to(idt), colidx[p0:p1].to(idt), vals[p0:p1],
                                    size=(r1 - r0, ncols))
        C = torch.sparse.mm(A, B)
        cr = C.crow_indices().to(torch.int64)
        crows.append(cr[:-1] + base if r1 < nrows else cr + base)
        cols.append(C.col_indices().to(torch.int32))
        vs.append(C.values())
        base += int(cr[-1].item())
        del A, C
    return torch.cat(crows).contiguous(), torch.cat(cols).contiguous(), torch.cat(vs).contiguous()


class Multigrid:
    """Device side: Galerkin operators, smoothers and the V-cycle."""

    def __init__(self, engine, tile=8, cells_per_block=64, coarsest_cells=64):
        self.e = engine
        self.lib = engine.lib
        self.dev = engine.dev
        self.nf = engine.nf
        mesh = engine.solver.mesh
        if engine.family != "DG":
            raise NotImplementedError("multigrid is implemented for DG spaces")
        # across ranks every rank builds the hierarchy of its own cells: the V-cycle is then the block solver of
        # PETSc's bjacobi (one block per rank, echemfem/solver.py:1186), couplings to ghost cells dropped
        self.distributed = engine.halo is not None
        if getattr(mesh, "tile_shape", None) is not None and len(set(mesh.tile_shape)) > 1:
            tile = mesh.tile_shape          # elongated tiles (blocks along the flow): coarse levels follow the mesh
            cells_per_block = int(np.prod(tile))
        self.h = Hierarchy(mesh, engine.solver.W.scalar.degree, tile=tile, coarsest_cells=coarsest_cells,
                           fine_rows=engine.N)
        if self.h.nlevels < 2:
            raise NotImplementedError("mesh cannot be coarsened")
        n = self.h.n
        self.lv = []
        for l, L in enumerate(self.h.levels):
            D = Level()
            D.N = L.ncell * n if l else engine.N  # field stride (finest level: owned + ghost cells)
            D.ndof = D.N * self.nf
            # one warp solves one block row by row: small blocks keep the smoother of the coarse levels (few
            # cells, little parallelism) from dominating the cycle; the fine level follows the mesh tiles
            nb = max(1, L.ncell // cells_per_block)
            if l == 0 and getattr(mesh, "block_hint", None):
                nb = max(1, L.ncell // int(mesh.block_hint))
            if os.environ.get("ECH_MG_CELLS_PER_BLOCK"):             # tuning hook (tools/newton_scale.py)
                nb = max(1, L.ncell // int(os.environ["ECH_MG_CELLS_PER_BLOCK"]))
            D.nblocks = nb
            D.block_edges = np.linspace(0, L.ncell, nb + 1).astype(np.int64)
            D.block_ptr = torch.from_numpy(D.block_edges).to(self.dev)
            D.ilu_plan = None
            D.r = torch.zeros(D.ndof, dtype=torch.float64, device=self.dev)
            D.z = torch.zeros(D.ndof, dtype=torch.float64, device=self.dev)
            D.t = torch.zeros(D.ndof, dtype=torch.float64, device=self.dev)
            D.s = torch.zeros(D.ndof, dtype=torch.float64, device=self.dev)
            self.lv.append(D)
        for l, P in enumerate(self.h.P):
            D = self.lv[l]
            D.P = _dev_csr(P, self.dev)
            D.R = _dev_csr(P.T, self.dev)
            # block-diagonal (all fields) transfer operators as torch CSR tensors for the Galerkin products
            if os.environ.get("ECH_MG_SPGEMM", "0") != "0" or n not in (2, 4, 8):
                # block-diagonal (all fields) transfer operators as torch CSR tensors for the cuSPARSE cross-check path
                D.Pt = self._blockdiag(P)
                D.Rt = self._blockdiag(P.T.tocsr())
        self.nlevels = len(self.lv)
        # own Galerkin kernel (ech_galerkin_dg): fixed block-stencil pattern and cell tables of every coarse level
        self.own_galerkin = os.environ.get("ECH_MG_SPGEMM", "0") == "0" and n in (2, 4, 8)
        if self.own_galerkin:
            self._galerkin_tables(engine)
        self.planned = os.environ.get("ECH_ILU_PLANNED", "1") != "0"
        # continuous-Q1 auxiliary space for the potential fields (auxspace.py: why the nested DG hierarchy alone is
        # a weak preconditioner for the elliptic block, and what this buys)
        self.aux = None
        if os.environ.get("ECH_MG_AUX", "1") != "0" and self.h.p == 1:
            s = engine.solver
            fields = [getattr(s, nm) for nm in ("i_Ul", "i_Us") if getattr(s, nm, None) is not None]
            if fields:
                from .auxspace import AuxSpaceCG
                self.aux = AuxSpaceCG(self, fields)
                if getattr(engine, "rowptr", None) is not None:
                    self.aux.prepare(engine.rowptr, engine.colidx)
        self.aux_overlap = os.environ.get("ECH_MG_AUX_OVERLAP", "1") != "0"
        self._side = None
        self.use_graph = os.environ.get("ECH_MG_GRAPH", "1") != "0"
        self._graph = self._gr = self._gz = None

    def _galerkin_tables(self, engine):
        from .mesh import TensorMesh
        dev, n, nf = self.dev, self.h.n, self.nf
        own = np.ascontiguousarray(engine.form.own_mask, dtype=bool)
        nbm = np.ascontiguousarray(engine.form.nbr_mask, dtype=bool)
        f = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)
        self._own_d, self._nbm_d = f(own.astype(np.uint8), torch.uint8), f(nbm.astype(np.uint8), torch.uint8)
        d = len(self.h.levels[0].shape)
        for l, (L, D) in enumerate(zip(self.h.levels, self.lv)):
            if l == 0:
                m = engine.m                      # the engine's own tables (ghost cells lie behind the owned ones)
                D.tab = {"nbr": m["nbr"], "slot": m["slot"], "slot_self": m["slot_self"], "ncol": m["ncol"]}
                D.ncell_owned = engine.ncell
            else:
                tile = self.h.levels[l - 1].ctile
                cm = TensorMesh(L.axes, tile=tile)
                nbr, _, slot, slot_self, ncol = cm.connectivity({})
                nbr = np.where(nbr >= 0, nbr, -1)
                cs, nc = _sorted_cells(nbr)
                rp, ci = dg_pattern(nf, n, L.ncell, cs, nc, own, nbm)
                D.tab = {"nbr": f(nbr, torch.int32), "slot": f(slot, torch.uint8),
                         "slot_self": f(slot_self, torch.uint8), "ncol": f(ncol, torch.uint8)}
                D.ncell_owned = L.ncell
                D.rowptr, D.colidx = f(rp, torch.int64), f(ci, torch.int32)
                D.vals = torch.zeros(int(rp[-1]), dtype=torch.float64, device=dev)
            if l < len(self.h.P):
                P = self.h.P[l].tocoo()
                pblk = np.zeros((L.ncell * n, n))
                keep = P.row < L.ncell * n
                pblk[P.row[keep], P.col[keep] % n] = P.data[keep]
                ncc = self.h.levels[l + 1].ncell
                children = -np.ones((ncc, 2 ** d), dtype=np.int32)
                order = np.argsort(L.parent, kind="stable")
                par = L.parent[order]
                first = np.searchsorted(par, np.arange(ncc))
                rank = np.arange(L.ncell) - first[par]
                children[par, rank] = order
                D.gal = {"parent": f(L.parent, torch.int32), "pblk": f(pblk, torch.float64),
                         "children": f(children, torch.int32), "nchild": 2 ** d}

    def _galerkin(self, D, Dc):
        """Dc.vals = P^T D.vals P (ech_galerkin_dg)."""
        n, nf, d = self.h.n, self.nf, len(self.h.levels[0].shape)
        ip = (capi.c_i64 * 8)(nf, n, 2 * d, D.gal["nchild"], D.ncell_owned, Dc.ncell_owned, D.N, Dc.N)
        arrs = [D.rowptr, D.vals, D.tab["nbr"], D.tab["slot"], D.tab["slot_self"], D.tab["ncol"], D.gal["parent"],
                D.gal["pblk"], D.gal["children"], Dc.rowptr, Dc.vals, Dc.tab["nbr"], Dc.tab["slot"],
                Dc.tab["slot_self"], Dc.tab["ncol"], self._own_d, self._nbm_d]
        pp = (capi.c_p * 17)(*[a.data_ptr() for a in arrs])
        capi.check(self.lib.ech_galerkin_dg(ip, pp, self.e._stream()))

    def _blockdiag(self, P):
        P = P.tocsr()
        P.sort_indices()
        nr, nc = P.shape
        nf = self.nf
        crow = np.concatenate([P.indptr[:-1] + f * P.nnz for f in range(nf)] + [[nf * P.nnz]])
        col = np.concatenate([P.indices + f * nc for f in range(nf)])
        val = np.tile(P.data, nf)
        it = np.int32 if max(nf * P.nnz, nf * nc) < 2 ** 31 - 1 else np.int64
        return torch.sparse_csr_tensor(torch.from_numpy(crow.astype(it)).to(self.dev),
                                       torch.from_numpy(col.astype(it)).to(self.dev),
                                       torch.from_numpy(val).to(self.dev), size=(nf * nr, nf * nc))

    # -- setup: Galerkin products + ILU factors, once per Newton iteration -------------------------------
    def setup(self):
        e, lib, st = self.e, self.lib, self.e._stream()
        fine = self.lv[0]
        fine.rowptr, fine.colidx, fine.vals = e.rowptr, e.colidx, e.vals
        if self.aux is not None:
            self.aux.setup(fine.rowptr, fine.colidx, fine.vals, st)
        for l in range(self.nlevels - 1):
            D, Dc = self.lv[l], self.lv[l + 1]
            if self.own_galerkin:
                self._galerkin(D, Dc)
                continue
            # cross-check path (ECH_MG_SPGEMM=1): A P, then P^T (A P) through cuSPARSE, each in row chunks
            ap = spgemm_rows(D.rowptr, D.colidx, D.vals, D.ndof, D.Pt, D.Pt.values().numel() / D.ndof)
            AP = torch.sparse_csr_tensor(ap[0].to(D.Pt.crow_indices().dtype), ap[1].to(D.Pt.crow_indices().dtype), ap[2],
                                         size=(D.ndof, Dc.ndof))
            Rt = D.Rt
            Dc.rowptr, Dc.colidx, Dc.vals = spgemm_rows(Rt.crow_indices().to(torch.int64), Rt.col_indices().to(torch.int32),
                                                        Rt.values(), D.ndof, AP, ap[2].numel() / D.ndof)
            del ap, AP
        for l, D in enumerate(self.lv[:-1]):
            if not hasattr(D, "lu") or D.lu.numel() != D.vals.numel():
                D.lu = torch.empty_like(D.vals)
                D.diag = torch.empty(D.ndof, dtype=torch.int32, device=self.dev)
            if self.planned and (D.ilu_plan is None or D.ilu_plan.lcol.numel() != D.vals.numel()):
                # level-scheduled solves and the factorisation on local columns: the plan follows the pattern
                # (fixed after the first Galerkin product)
                from .ilu import IluPlan
                D.ilu_plan = IluPlan(lib, self.dev, st, D.ndof, self.nf, D.N, self.h.n, D.block_edges, D.block_ptr,
                                     D.rowptr, D.colidx, int(D.vals.numel()))
                if not D.ilu_plan.ok:
                    D.ilu_plan = None
                    self.planned = False
        self._factor_levels()
        C = self.lv[-1]
        dense = torch.sparse_csr_tensor(C.rowptr, C.colidx.to(torch.int64), C.vals, size=(C.ndof, C.ndof)).to_dense()
        if getattr(C, "inv", None) is not None and C.inv.shape == dense.shape:
            C.inv.copy_(torch.linalg.inv(dense))
        else:
            C.inv = torch.linalg.inv(dense).contiguous()
        # the captured cycle reads fixed buffers: it stays valid while none of them moved (instantiating a graph of
        # ~250 kernels costs 30 - 650 ms, measured, against 12.8 ms per replay -- re-capturing per Newton iteration
        # was up to a quarter of a solve)
        sig = self._pointers()
        if sig != getattr(self, "_graph_sig", None):
            self._graph = None
            self._graph_sig = sig

    def _factor_levels(self):
        """ILU(0) of all levels at once, one stream per level (`ECH_MG_CONCURRENT_FACTOR=0`: one after the other on
        the current stream): the factorisation of a level is one warp per tile walking its rows in sequence, so the
        small levels are pure latency and hide behind the fine level.  With a plan the factorisation runs on the
        plan's local columns (`ech_bjacobi_ilu0_factor_planned`; `ECH_ILU_FACTOR_PLANNED=0`: the search-based kernel)."""
        lib = self.lib
        main = torch.cuda.current_stream()
        nl = len(self.lv) - 1
        concurrent = os.environ.get("ECH_MG_CONCURRENT_FACTOR", "1") != "0"
        planned = os.environ.get("ECH_ILU_FACTOR_PLANNED", "1") != "0"
        if getattr(self, "_fstreams", None) is None or len(self._fstreams) != nl:
            self._fstreams = [torch.cuda.Stream(device=self.dev) for _ in range(nl)]
            self._fflags = torch.zeros(nl, dtype=torch.int32, device=self.dev)
        ev = torch.cuda.Event()
        ev.record(main)
        es = self._fflags.element_size()
        for l, D in enumerate(self.lv[:-1]):
            sl = self._fstreams[l] if concurrent else main
            if concurrent:
                sl.wait_event(ev)
            st = capi.C.c_void_p(sl.cuda_stream)
            flag = capi.C.c_void_p(self._fflags.data_ptr() + l * es)
            if D.ilu_plan is not None and planned:
                D.ilu_plan.factor_async(D.rowptr, D.colidx, D.vals, D.lu, D.diag, flag, st)
                continue
            capi.check(lib.ech_bjacobi_ilu0_factor_async(D.ndof, self.nf, D.N, self.h.n, D.nblocks, ptr(D.block_ptr),
                                                         ptr(D.rowptr), ptr(D.colidx), ptr(D.vals), int(D.vals.numel()),
                                                         ptr(D.lu), ptr(D.diag), flag, st))
            if D.ilu_plan is not None:
                D.ilu_plan.refresh(D.lu, st)
        if concurrent:
            for sl in self._fstreams:
                main.wait_stream(sl)
        for flag in self._fflags.cpu().tolist():
            capi.check(lib.ech_bjacobi_ilu0_factor_status(int(flag)))

    def _pointers(self):
        out = []
        for D in self.lv:
            for name in ("rowptr", "colidx", "vals", "lu", "diag", "inv"):
                t = getattr(D, name, None)
                if t is not None:
                    out.append(t.data_ptr())
            if getattr(D, "ilu_plan", None) is not None:
                out.append(id(D.ilu_plan))
        if self.aux is not None:
            out += self.aux.h.pointers()
        return out

    # -- apply ---------------------------------------------------------------------------------------------
    def _smooth(self, D, r, z):
        if D.ilu_plan is not None:
            D.ilu_plan.apply(D.lu, r, z, self.e._stream())
            return
        capi.check(self.lib.ech_bjacobi_ilu0_apply(D.ndof, self.nf, D.N, self.h.n, D.nblocks, ptr(D.block_ptr),
                                                   ptr(D.rowptr), ptr(D.colidx), ptr(D.lu), ptr(D.diag), ptr(r), ptr(z),
                                                   self.e._stream()))

    def _spmv(self, D, x, y):
        capi.check(self.lib.ech_spmv(D.ndof, ptr(D.rowptr), ptr(D.colidx), ptr(D.vals), ptr(x), ptr(y), self.e._stream()))

    def _transfer(self, M, nrow, x, xstride, y, ystride):
        """y_f = M x_f for every field f (scalar operator applied field by field)."""
        rp, ci, v = M
        es = x.element_size()
        for f in range(self.nf):
            capi.check(self.lib.ech_spmv(nrow, ptr(rp), ptr(ci), ptr(v),
                                         capi.C.c_void_p(x.data_ptr() + f * xstride * es),
                                         capi.C.c_void_p(y.data_ptr() + f * ystride * es), self.e._stream()))

    def _restrict(self, D, Dc, rf, rc):
        """rc = P^T rf (all fields)."""
        if hasattr(D, "gal"):
            g = D.gal
            capi.check(self.lib.ech_dg_restrict(self.h.n, self.nf, g["nchild"], Dc.ncell_owned, D.N, Dc.N,
                                                ptr(g["children"]), ptr(g["pblk"]), ptr(rf), ptr(rc), self.e._stream()))
        else:
            self._transfer(D.R, Dc.N, rf, D.N, rc, Dc.N)

    def _prolong_add(self, D, Dc, zc, z):
        """z += P zc (all fields; rows behind the owned cells untouched)."""
        if hasattr(D, "gal"):
            g = D.gal
            capi.check(self.lib.ech_dg_prolong_add(self.h.n, self.nf, D.ncell_owned, D.N, Dc.N, ptr(g["parent"]),
                                                   ptr(g["pblk"]), ptr(zc), ptr(z), self.e._stream()))
        else:
            self._transfer(D.P, D.N, zc, Dc.N, D.t, D.N)
            capi.check(self.lib.ech_axpby(D.ndof, 1.0, ptr(D.t), 1.0, ptr(z), self.e._stream()))

    def vcycle(self, l, r, z):
        """z = V-cycle(r) on level l; r is not modified."""
        lib, st = self.lib, self.e._stream()
        D = self.lv[l]
        if l == self.nlevels - 1:
            capi.check(lib.ech_dense_gemv(D.ndof, D.ndof, ptr(D.inv), ptr(r), ptr(z), st))
            return
        Dc = self.lv[l + 1]
        self._smooth(D, r, z)                                                  # pre-smoothing
        self._spmv(D, z, D.t)
        capi.check(lib.ech_axpby(D.ndof, 1.0, ptr(r), -1.0, ptr(D.t), st))      # t = r - A z
        aux = self.aux if l == 0 else None
        side = None
        if aux is not None:
            # conforming correction from the same residual.  It is independent of the DG coarse-grid correction and
            # both are chains of small latency-bound kernels: run it on a second stream (a parallel branch of the
            # captured graph) unless the CSR transfer path, which overwrites D.t, is in use
            if self.aux_overlap and hasattr(D, "gal"):
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.dev)
                side = self._side
                side.wait_stream(torch.cuda.current_stream())
                aux.restrict_and_solve(D.t, capi.C.c_void_p(side.cuda_stream))
            else:
                aux.restrict_and_solve(D.t, st)
        self._restrict(D, Dc, D.t, Dc.r)
        self.vcycle(l + 1, Dc.r, Dc.z)
        self._prolong_add(D, Dc, Dc.z, z)                                       # z += P z_c
        if aux is not None:
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)
            aux.prolong_add(z, D.s, st)                                         # z_phi += Q z_aux
        self._spmv(D, z, D.t)
        capi.check(lib.ech_axpby(D.ndof, 1.0, ptr(r), -1.0, ptr(D.t), st))      # t = r - A z
        self._smooth(D, D.t, D.s)                                               # post-smoothing
        capi.check(lib.ech_axpby(D.ndof, 1.0, ptr(D.s), 1.0, ptr(z), st))

    def _apply_eager(self, r, z):
        if self.distributed:
            z.zero_()          # ghost entries (stale halo values) must not enter the rank-local residuals
        self.vcycle(0, r, z)

    def apply(self, r, z):
        """z = V-cycle(r).  The cycle is a fixed sequence of ~100 small launches on fixed buffers: it is captured
        once per set-up into a CUDA graph and replayed (the launch overhead of the Python-driven sequence is
        otherwise ~2 ms of a 16 ms cycle)."""
        if not self.use_graph:
            return self._apply_eager(r, z)
        if self._graph is None:
            if self._gr is None:
                self._gr, self._gz = torch.empty_like(r), torch.zeros_like(z)
            self._gr.copy_(r)
            self._apply_eager(self._gr, self._gz)          # warm-up outside the capture (function attributes)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._apply_eager(self._gr, self._gz)
            self._graph = g
        self._gr.copy_(r)
        self._graph.replay()
        z.copy_(self._gz)
