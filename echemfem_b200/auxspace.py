"""Continuous-Q1 auxiliary space for the potential fields of the DG multigrid preconditioner.

Why.  The Galerkin coarse operators of nested DG spaces inherit the interior-penalty term of the FINE mesh
(sigma / h_fine on every coarse face), so on level l the jump penalty is 2^l times too strong for that level and the
smoothers mostly see jump modes: the V-cycle of `multigrid.py` is not h-independent on the elliptic (potential)
block.  On the device, cfg2 physics: 27 / 39 / 77 / 130 GMRES iterations per solve at N = 128 / 256 / 512 / 1408.
A scipy prototype of the same cycle on oracle Jacobians (tests/mg_prototype.py, random right-hand side, rtol 1e-12)
shows the same growth, 23 / 30 / 45 / 69 at N = 32 / 64 / 128 / 256, the potential block alone 28 / 36 at N = 64 / 128, the
(advective) concentration block alone 12 / 22.  A CONTINUOUS coarse function has no jumps, so the Galerkin
operator of the conforming subspace carries no penalty at all and its multigrid hierarchy (bilinear interpolation,
damped Jacobi) is the textbook one.  The correction
        z_phi += Q V_cg(Q^T r_phi),      Q: continuous Q1 vertex values -> values at the DG nodes (the embedding),
is added at the finest level next to the DG coarse-grid correction (same residual: no extra fine-level SpMV).  With it
the prototype needs 25 / 27 / 32 / 40 iterations at N = 32 / 64 / 128 / 256 and the device 58 / 61 instead of
130 / 140 at N = 1408.  Only the potential fields get it: for the advection-dominated concentration field the
conforming Galerkin operator is the central (unstable) discretisation and GMRES no longer converges (prototype: 300+
iterations; tests/test_mg_prototype.py), their nested DG hierarchy keeps the upwinding.  Rediscretising the coarse
DG levels instead of Galerkin products is worse (47 / 64 against 30 / 45).

The reference reaches the same kind of hierarchy through PETSc (`gmg`: geometric multigrid on a MeshHierarchy, `amg`:
hypre on the assembled operator, echemfem/solver.py:1190-1213); this is the hand-built equivalent for tensor meshes.

Everything in the apply path is this package's own kernels (`ech_spmv`, `ech_axpby`, `ech_dense_gemv`,
`ech_aux_prolong_add`); the set-up uses `ech_aux_galerkin` for Q^T A Q (cell block by cell block, fixed pattern) and
cuSPARSE SpGEMM through `torch.sparse.mm` for the small conforming coarse levels (<= 1/16 of the fine matrix in total).
"""
import warnings

import numpy as np
import scipy.sparse as sp
import torch

# torch announces on first use that sparse CSR tensors are a beta feature and that invariant checks are off; the
# matrices here are built from sorted, duplicate-free scipy CSR data
warnings.filterwarnings("ignore", message=".*Sparse CSR tensor support is in beta.*")
warnings.filterwarnings("ignore", message=".*Sparse invariant checks are implicitly disabled.*")

from . import capi
from .capi import ptr


def interpolation_1d(ax_f, ax_c):
    """(len(ax_f) x len(ax_c)) linear interpolation between the vertex sets of one direction (by coordinate, so
    graded axes and the odd-count coarsening of `multigrid.Hierarchy` are covered)."""
    ax_f, ax_c = np.asarray(ax_f, dtype=float), np.asarray(ax_c, dtype=float)
    j = np.clip(np.searchsorted(ax_c, ax_f, side="right") - 1, 0, len(ax_c) - 2)
    t = (ax_f - ax_c[j]) / (ax_c[j + 1] - ax_c[j])
    t = np.where(np.abs(t) < 1e-12, 0.0, np.where(np.abs(t - 1.0) < 1e-12, 1.0, t))
    i = np.arange(len(ax_f))
    P = sp.csr_matrix((np.concatenate([1.0 - t, t]), (np.concatenate([i, i]), np.concatenate([j, j + 1]))),
                      shape=(len(ax_f), len(ax_c)))
    P.eliminate_zeros()
    return P


def dg_to_cg_nodes(shape, inv_perm, nrows):
    """Vertex table of a DG-Q1 space: entry (cell number * 2^d + v) is the CG vertex number (lexicographic over the
    (s_k + 1) vertices per direction) of corner v of the cell, v lexicographic like the DG nodes (the layout of
    `multigrid._dof_map`); -1 for rows past the owned cells."""
    d = len(shape)
    sizes = [2 * s for s in shape]
    idx = np.arange(int(np.prod(sizes)), dtype=np.int64)
    parts = []
    rem = idx
    for k in range(d - 1, -1, -1):
        parts.append(rem % sizes[k])
        rem = rem // sizes[k]
    parts = parts[::-1]
    cell = np.zeros_like(idx)
    node = np.zeros_like(idx)
    vert = np.zeros_like(idx)
    for k in range(d):
        cell = cell * shape[k] + parts[k] // 2
        node = node * 2 + parts[k] % 2
        vert = vert * (shape[k] + 1) + parts[k] // 2 + parts[k] % 2
    if inv_perm is not None:
        cell = np.asarray(inv_perm, dtype=np.int64)[cell]
    out = np.full(int(nrows), -1, dtype=np.int64)
    out[cell * (2 ** d) + node] = vert
    return out


def embedding_block(d):
    """W[a][v] = value of the continuous hat function of corner v at DG node a of the reference cell (both
    lexicographic): the embedding Q restricted to one cell, identical for all cells of a tensor mesh."""
    from .tables import element_nodes
    xi = np.asarray(element_nodes(1, "DG"), dtype=float)
    w1 = np.stack([1.0 - xi, xi], axis=1)                 # [node][corner] in one direction
    W = w1
    for _ in range(d - 1):
        W = np.kron(W, w1)
    return np.ascontiguousarray(W)


class _Lv:
    pass


def block_heads(rowptr, colidx, N, nv, fields, dev, chunk_rows=1 << 20):
    """Cell blocks of a field-blocked DG matrix that couple two of `fields`: a block is found by its first entry
    (row node 0, column node 0).  Returns device tensors (row0, off, row cell, col cell, ia, ja): first row of the
    block, offset of the block from each of its row starts, the two cell numbers and the positions of the two
    fields in `fields`."""
    out = [[] for _ in range(6)]
    for ia, fa in enumerate(fields):
        for r0 in range(fa * N, (fa + 1) * N, chunk_rows):
            r1 = min((fa + 1) * N, r0 + chunk_rows)
            p0, p1 = int(rowptr[r0].item()), int(rowptr[r1].item())
            if p1 == p0:
                continue
            counts = rowptr[r0 + 1:r1 + 1] - rowptr[r0:r1]
            rows = torch.repeat_interleave(torch.arange(r0, r1, device=dev), counts)
            cols = colidx[p0:p1].to(torch.int64)
            cf = cols // N
            ja = torch.full_like(cf, -1)
            for k, fb in enumerate(fields):
                ja[cf == fb] = k
            rl, cl = rows - fa * N, cols - cf * N
            pos = torch.nonzero((ja >= 0) & (rl % nv == 0) & (cl % nv == 0)).reshape(-1)
            rsel = rows[pos]
            out[0].append(rsel)
            out[1].append((pos + p0 - rowptr[rsel]).to(torch.int32))
            out[2].append(rl[pos] // nv)
            out[3].append(cl[pos] // nv)
            out[4].append(torch.full_like(rsel, ia))
            out[5].append(ja[pos])
            del rows, cols, cf, ja, rl, cl, pos, rsel
    return [torch.cat(o) if o else torch.zeros(0, dtype=torch.int64, device=dev) for o in out]


def coarse_embedding(x, nv, cax):
    """Embedding of the continuous Q1 functions on the vertex grid `cax` (one coordinate array per direction) into a
    DG-Q1 space with node coordinates x (N, d), nv nodes per cell, cells contiguous: per cell the multi-index of
    the coarse cell that contains it (ncell, d); per node the values of the 2^d corner functions of that coarse
    cell (N, 2^d, corners lexicographic) and their vertex numbers (N, 2^d).  Coordinates only, so every rank of
    a partitioned mesh numbers the shared vertices alike."""
    N, d = x.shape
    I = np.zeros((N, d), dtype=np.int64)
    w = np.zeros((N, d, 2))
    for k in range(d):
        I[:, k] = np.clip(np.searchsorted(cax[k], x[:, k], side="right") - 1, 0, len(cax[k]) - 2)
        t = (x[:, k] - cax[k][I[:, k]]) / (cax[k][I[:, k] + 1] - cax[k][I[:, k]])
        w[:, k, 0], w[:, k, 1] = 1.0 - t, t
    cI = I.reshape(N // nv, nv, d)
    if not np.all(cI == cI[:, :1, :]):
        raise NotImplementedError("fine cells straddle coarse cells: the coarse grid must be nested in the mesh")
    vdims = [len(a) for a in cax]
    W = np.ones((N, 2 ** d))
    vert = np.zeros((N, 2 ** d), dtype=np.int64)
    for v in range(2 ** d):
        lin = np.zeros(N, dtype=np.int64)
        for k in range(d):
            b = (v >> (d - 1 - k)) & 1
            W[:, v] *= w[:, k, b]
            lin = lin * vdims[k] + I[:, k] + b
        vert[:, v] = lin
    return cI[:, 0, :], W, vert


class ConformingHierarchy:
    """Multigrid for a continuous-Q1 operator on the vertices of a tensor mesh: bilinear interpolation between
    vertex sets (every other vertex dropped per level), Galerkin coarse operators, damped Jacobi, dense coarsest
    level.  Level 0's CSR operator (`lv[0].rowptr / colidx / vals`, `na` fields blocked by field) is the caller's."""

    def __init__(self, lib, dev, axes, na, omega=0.8, coarsest_nodes=400):
        self.lib, self.dev, self.na, self.omega = lib, dev, int(na), float(omega)
        axes = [np.asarray(a, dtype=float) for a in axes]
        self.lv = []
        while True:
            D = _Lv()
            D.nnode = int(np.prod([len(a) for a in axes]))
            D.n = D.nnode * self.na
            for name in ("r", "z", "t", "s"):
                setattr(D, name, torch.zeros(D.n, dtype=torch.float64, device=dev))
            self.lv.append(D)
            if D.nnode <= coarsest_nodes or any(len(a) < 4 for a in axes) or len(self.lv) >= 14:
                break
            cax = [a[::2] if (len(a) - 1) % 2 == 0 else np.append(a[::2], a[-1]) for a in axes]
            P = interpolation_1d(axes[0], cax[0])
            for k in range(1, len(axes)):
                P = sp.kron(P, interpolation_1d(axes[k], cax[k]), format="csr")
            D.P = _dev_csr(P, dev)
            D.R = _dev_csr(P.T, dev)
            Pb = sp.block_diag([P] * self.na, format="csr") if self.na > 1 else P.tocsr()
            D.Pt = _torch_csr(Pb, dev)
            D.Rt = _torch_csr(Pb.T.tocsr(), dev)
            axes = cax

    def setup_levels(self):
        """Coarse operators, smoothers and the coarsest inverse from the current level-0 values."""
        for l in range(len(self.lv) - 1):
            D, Dc = self.lv[l], self.lv[l + 1]
            idt = D.Pt.crow_indices().dtype
            At = torch.sparse_csr_tensor(D.rowptr.to(idt), D.colidx.to(idt), D.vals, size=(D.n, D.n))
            Ac = torch.sparse.mm(D.Rt, torch.sparse.mm(At, D.Pt))
            if getattr(Dc, "vals", None) is not None and Dc.vals.numel() == Ac.values().numel():
                # same pattern as at the last set-up: update in place, so that a captured V-cycle stays valid
                Dc.rowptr.copy_(Ac.crow_indices())
                Dc.colidx.copy_(Ac.col_indices())
                Dc.vals.copy_(Ac.values())
            else:
                Dc.rowptr = Ac.crow_indices().to(torch.int64).contiguous()
                Dc.colidx = Ac.col_indices().to(torch.int32).contiguous()
                Dc.vals = Ac.values().contiguous()
            del At, Ac
        for D in self.lv[:-1]:
            # damped Jacobi as a diagonal CSR matrix: the smoother is one ech_spmv
            counts = D.rowptr[1:] - D.rowptr[:-1]
            rows = torch.repeat_interleave(torch.arange(D.n, device=self.dev), counts)
            diag = torch.zeros(D.n, dtype=torch.float64, device=self.dev)
            sel = D.colidx.to(torch.int64) == rows
            diag[rows[sel]] = D.vals[sel]
            if bool((diag == 0).any()):
                raise RuntimeError("auxiliary space: zero diagonal entry in a conforming level operator")
            if getattr(D, "w", None) is not None and D.w.numel() == D.n:
                torch.div(self.omega, diag, out=D.w)
            else:
                D.w = (self.omega / diag).contiguous()
            if not hasattr(D, "wrow") or D.wrow.numel() != D.n + 1:
                D.wrow = torch.arange(D.n + 1, dtype=torch.int64, device=self.dev)
                D.wcol = torch.arange(D.n, dtype=torch.int32, device=self.dev)
        C = self.lv[-1]
        dense = torch.sparse_csr_tensor(C.rowptr, C.colidx.to(torch.int64), C.vals, size=(C.n, C.n)).to_dense()
        if getattr(C, "inv", None) is not None and C.inv.shape == dense.shape:
            C.inv.copy_(torch.linalg.inv(dense))
        else:
            C.inv = torch.linalg.inv(dense).contiguous()

    def pointers(self):
        """Addresses of everything a captured cycle reads (a CUDA graph stays valid while these do not move)."""
        out = []
        for D in self.lv:
            for name in ("rowptr", "colidx", "vals", "w", "inv"):
                t = getattr(D, name, None)
                if t is not None:
                    out.append(t.data_ptr())
        return out

    def _spmv(self, rp, ci, v, n, x, y, st):
        capi.check(self.lib.ech_spmv(n, ptr(rp), ptr(ci), ptr(v), ptr(x), ptr(y), st))

    def transfer(self, M, nrow, x, xstride, y, ystride, st):
        """y_a = M x_a for every field a (scalar operator, vectors blocked by field with the given strides)."""
        rp, ci, v = M
        es = x.element_size()
        for a in range(self.na):
            capi.check(self.lib.ech_spmv(nrow, ptr(rp), ptr(ci), ptr(v), capi.C.c_void_p(x.data_ptr() + a * xstride * es),
                                         capi.C.c_void_p(y.data_ptr() + a * ystride * es), st))

    def cycle(self, l, st):
        """lv[l].z = V(1,1)(lv[l].r)."""
        lib = self.lib
        D = self.lv[l]
        if l == len(self.lv) - 1:
            capi.check(lib.ech_dense_gemv(D.n, D.n, ptr(D.inv), ptr(D.r), ptr(D.z), st))
            return
        Dc = self.lv[l + 1]
        self._spmv(D.wrow, D.wcol, D.w, D.n, D.r, D.z, st)                       # z = omega D^-1 r
        self._spmv(D.rowptr, D.colidx, D.vals, D.n, D.z, D.t, st)
        capi.check(lib.ech_axpby(D.n, 1.0, ptr(D.r), -1.0, ptr(D.t), st))         # t = r - A z
        self.transfer(D.R, Dc.nnode, D.t, D.nnode, Dc.r, Dc.nnode, st)
        self.cycle(l + 1, st)
        self.transfer(D.P, D.nnode, Dc.z, Dc.nnode, D.t, D.nnode, st)
        capi.check(lib.ech_axpby(D.n, 1.0, ptr(D.t), 1.0, ptr(D.z), st))          # z += P z_c
        self._spmv(D.rowptr, D.colidx, D.vals, D.n, D.z, D.t, st)
        capi.check(lib.ech_axpby(D.n, 1.0, ptr(D.r), -1.0, ptr(D.t), st))         # t = r - A z
        self._spmv(D.wrow, D.wcol, D.w, D.n, D.t, D.s, st)
        capi.check(lib.ech_axpby(D.n, 1.0, ptr(D.s), 1.0, ptr(D.z), st))          # z += omega D^-1 t


class AuxSpaceCG:
    """The conforming space on the cells of ONE rank (all cells on one GPU), applied inside the V-cycle."""

    def __init__(self, mg, fields, omega=0.8, coarsest_nodes=400):
        self.mg, self.lib, self.dev = mg, mg.lib, mg.dev
        self.fields = [int(f) for f in fields]
        self.na = len(self.fields)
        L0 = mg.h.levels[0]
        if mg.h.p != 1:
            raise NotImplementedError("auxiliary space: Q1 only")
        self.N = int(mg.lv[0].N)                         # rows per field of the fine level (owned + ghost)
        axes = [np.asarray(a, dtype=float) for a in L0.axes]
        shape = tuple(len(a) - 1 for a in axes)
        self.nv = 2 ** len(shape)
        vtab = dg_to_cg_nodes(shape, L0.inv, self.N)       # (cell, corner) -> vertex
        self.W = embedding_block(len(shape))
        self.ncg = int(np.prod([s + 1 for s in shape]))
        owned = np.nonzero(vtab >= 0)[0]
        cell, a = owned // self.nv, owned % self.nv
        rows = np.repeat(owned, self.nv)
        cols = vtab.reshape(-1, self.nv)[cell].reshape(-1)
        Q = sp.csr_matrix((self.W[a].reshape(-1), (rows, cols)), shape=(self.N, self.ncg))
        self.Q = _dev_csr(Q, self.dev)
        self.Qt = _dev_csr(Q.T, self.dev)
        self.vtab = torch.from_numpy(vtab).to(self.dev)
        self.Wd = torch.from_numpy(self.W.reshape(-1)).to(self.dev)
        self.h = ConformingHierarchy(self.lib, self.dev, axes, self.na, omega, coarsest_nodes)
        self.lv = self.h.lv
        self._pattern = None

    # -- set-up --------------------------------------------------------------------------------------------------
    def _symbolic(self, rowptr, colidx):
        """Blocks on owned cells, and where their W^T B W lands in A_aux."""
        ncg, na, nv = self.ncg, self.na, self.nv
        nrow = na * ncg
        vt = self.vtab.view(-1, nv)
        row0, off, rc, cc, ia, ja = block_heads(rowptr, colidx, self.N, nv, self.fields, self.dev)
        vi, vj = vt[rc], vt[cc]                                              # (blocks, nv) vertex numbers
        ok = (vi[:, 0] >= 0) & (vj[:, 0] >= 0)                               # owned cells on both sides
        row0, off, vi, vj, ia, ja = row0[ok], off[ok], vi[ok], vj[ok], ia[ok], ja[ok]
        keys = ((ia[:, None] * ncg + vi)[:, :, None] * nrow + (ja * ncg)[:, None, None] + vj[:, None, :]).reshape(-1)
        del vi, vj, rc, cc, ok
        uniq, inv = torch.unique(keys, return_inverse=True)
        del keys
        A = self.lv[0]
        A.colidx = (uniq % nrow).to(torch.int32).contiguous()
        rcount = torch.bincount(uniq // nrow, minlength=nrow)
        A.rowptr = torch.zeros(nrow + 1, dtype=torch.int64, device=self.dev)
        A.rowptr[1:] = torch.cumsum(rcount, 0)
        A.vals = torch.zeros(uniq.numel(), dtype=torch.float64, device=self.dev)
        self.row0 = row0.contiguous()
        self.off = off.contiguous()
        self.dst = inv.to(torch.int32).contiguous()
        self._pattern = (int(rowptr.data_ptr()), int(colidx.numel()))

    def prepare(self, rowptr, colidx):
        """Pattern-only part of the set-up (block lists and the pattern of A_aux): done once per sparsity pattern,
        with the hierarchy, like the tables of the DG Galerkin kernel."""
        if self._pattern != (int(rowptr.data_ptr()), int(colidx.numel())):
            self._symbolic(rowptr, colidx)

    def setup(self, rowptr, colidx, vals, stream):
        """Level operators for the current fine matrix (called from Multigrid.setup, once per Jacobian)."""
        self.prepare(rowptr, colidx)
        A = self.lv[0]
        capi.check(self.lib.ech_aux_galerkin(self.nv, self.row0.numel(), ptr(self.row0), ptr(self.off), ptr(self.dst),
                                             ptr(rowptr), ptr(vals), ptr(self.Wd), A.vals.numel(), ptr(A.vals), stream))
        self.h.setup_levels()

    # -- apply ---------------------------------------------------------------------------------------------------
    def restrict_and_solve(self, res, st):
        """One V-cycle of the auxiliary problem for the fine-level residual `res` (all fields, stride N)."""
        D = self.lv[0]
        es = res.element_size()
        rp, ci, v = self.Qt
        for a, f in enumerate(self.fields):
            capi.check(self.lib.ech_spmv(self.ncg, ptr(rp), ptr(ci), ptr(v), capi.C.c_void_p(res.data_ptr() + f * self.N * es),
                                         capi.C.c_void_p(D.r.data_ptr() + a * self.ncg * es), st))
        self.h.cycle(0, st)

    def prolong_add(self, z, tmp, st):
        """z_f += Q z_aux for every auxiliary field f; tmp: scratch of at least N entries."""
        D = self.lv[0]
        es = z.element_size()
        for a, f in enumerate(self.fields):
            capi.check(self.lib.ech_aux_prolong_add(self.nv, self.N, ptr(self.vtab), ptr(self.Wd),
                                                    capi.C.c_void_p(D.z.data_ptr() + a * self.ncg * es),
                                                    capi.C.c_void_p(z.data_ptr() + f * self.N * es), st))


class GlobalConformingCoarse:
    """Coarse space shared by ALL ranks: continuous Q1 on the global tensor mesh coarsened `2^level` times per
    direction, for the potential fields.  Block Jacobi across ranks (one V-cycle per rank, PETSc's bjacobi) cannot
    move the smooth global modes of the potential equation, and the interface couplings it drops act like a
    Dirichlet condition on every cut.  Here the Galerkin operator E^T A E is formed from the COMPLETE owned rows
    (ghost columns included, so the cut facets keep their consistency terms), summed over the ranks (one
    all-reduce of its values per set-up), and every rank runs the same small multigrid on it redundantly (one
    all-reduce of the restricted residual per Krylov iteration); E embeds the coarse vertex functions into the DG
    space by evaluating them at the DG nodes.

    Measured, 2 GPUs, 2 x 1408^2 cells (cfg2 physics, ksp_rtol 1e-12): box aggregation (4096 constants) 416 Krylov
    iterations; this space coarsened 8x / 4x / 2x / 1x per direction: 125 / 100 / 87 / 87 (one GPU, same physics on
    1408^2: 58).  Default 2x (level 1), coarsened further while it has more than 1.2 M vertices (4x at 4 and 8 GPUs of the
    weak-scaling benchmark)."""

    S1 = 5                    # vertex offsets -2..2 per direction between vertices of face-adjacent coarse cells

    def __init__(self, engine, fields, global_axes, level=None, omega=0.8, max_vertices=None):
        import os
        import torch.distributed as dist
        if level is None:
            level = int(os.environ.get("ECH_GCOARSE_LEVEL", "1"))
        if max_vertices is None:
            max_vertices = int(os.environ.get("ECH_GCOARSE_MAX_VERTICES", "1200000"))
        self.e, self.lib, self.dev = engine, engine.lib, engine.dev
        self.fields = [int(f) for f in fields]
        self.na = len(self.fields)
        self.N = int(engine.N)
        self.nv = int(engine.nloc)
        gax = [np.asarray(a, dtype=float) for a in global_axes]
        d = self.d = len(gax)
        if self.nv != 2 ** d:
            raise NotImplementedError("global conforming coarse space: Q1 only")
        while level > 0 and any((len(a) - 1) % (2 ** level) for a in gax):
            level -= 1
        while np.prod([(len(a) - 1) // 2 ** level + 1 for a in gax]) > max_vertices and \
                not any((len(a) - 1) % (2 ** (level + 1)) for a in gax):
            level += 1
        self.level = level
        cax = [a[::2 ** level] for a in gax]
        self.vdims = [len(a) for a in cax]
        self.nvert = int(np.prod(self.vdims))
        self.S = self.S1 ** d
        # per local cell (owned + ghost): coarse cell multi-index and the nv x nv embedding block
        x = np.asarray(engine.solver.W.scalar.node_coords(), dtype=float)            # (N, d)
        assert x.shape[0] == self.N
        cI, W, vert = coarse_embedding(x, self.nv, cax)
        self.cI = torch.from_numpy(np.ascontiguousarray(cI).astype(np.int32)).to(self.dev)
        self.Wc = torch.from_numpy(np.ascontiguousarray(W)).to(self.dev)              # [cell][node][corner]
        nown = engine.ncell * self.nv
        rows = np.repeat(np.arange(nown), self.nv)
        E = sp.csr_matrix((W[:nown].reshape(-1), (rows, vert[:nown].reshape(-1))), shape=(self.N, self.nvert))
        self.E = _dev_csr(E, self.dev)
        self.Et = _dev_csr(E.T, self.dev)
        self.fidx = torch.full((engine.nf,), -1, dtype=torch.int32, device=self.dev)
        for a, f in enumerate(self.fields):
            self.fidx[f] = a
        self.vd = torch.tensor(self.vdims, dtype=torch.int32, device=self.dev)
        # ELL layout of E^T A E (na * S entries per row) -> CSR of the in-range entries
        nrow = self.na * self.nvert
        self.ell = torch.zeros(nrow * self.na * self.S, dtype=torch.float64, device=self.dev)
        multi = np.stack(np.unravel_index(np.arange(self.nvert), self.vdims), axis=1)           # (nvert, d)
        offs = np.stack(np.unravel_index(np.arange(self.S), (self.S1,) * d), axis=1) - 2          # (S, d)
        J = multi[:, None, :] + offs[None, :, :]
        valid = np.all((J >= 0) & (J < np.asarray(self.vdims)[None, None, :]), axis=2)           # (nvert, S)
        Jlin = np.ravel_multi_index(tuple(np.where(valid, J[:, :, k], 0) for k in range(d)), self.vdims)
        sel, cols, counts = [], [], []
        for ia in range(self.na):
            vmask = np.tile(valid, (1, self.na))                                                 # (nvert, na * S)
            cc = np.concatenate([ja * self.nvert + Jlin for ja in range(self.na)], axis=1)
            base = (ia * self.nvert + np.arange(self.nvert))[:, None] * (self.na * self.S) + np.arange(self.na * self.S)[None, :]
            sel.append(base[vmask])
            cols.append(cc[vmask])
            counts.append(vmask.sum(axis=1))
        self.sel = torch.from_numpy(np.concatenate(sel)).to(self.dev)
        self.h = ConformingHierarchy(self.lib, self.dev, cax, self.na, omega)
        A = self.h.lv[0]
        A.colidx = torch.from_numpy(np.concatenate(cols).astype(np.int32)).to(self.dev)
        A.rowptr = torch.from_numpy(np.concatenate([[0], np.cumsum(np.concatenate(counts))]).astype(np.int64)).to(self.dev)
        self.res = torch.empty(engine.ndof, dtype=torch.float64, device=self.dev)
        self.tmp = torch.empty(self.N, dtype=torch.float64, device=self.dev)
        self.dist = dist
        self._pattern = None

    def local_operator(self, stream):
        """This rank's rows of E^T A E on the CSR pattern of level 0 (before the sum over ranks)."""
        e = self.e
        key = (int(e.rowptr.data_ptr()), int(e.colidx.numel()))
        if self._pattern != key:
            row0, off, rc, cc, ia, ja = block_heads(e.rowptr, e.colidx, self.N, self.nv, self.fields, self.dev)
            own = rc < e.ncell
            self.row0, self.off = row0[own].contiguous(), off[own].contiguous()
            self._pattern = key
        capi.check(self.lib.ech_aux_galerkin_global(self.nv, self.d, self.row0.numel(), ptr(self.row0), ptr(self.off),
                                                    ptr(e.rowptr), ptr(e.colidx), ptr(e.vals), self.N, ptr(self.fidx),
                                                    self.na, ptr(self.Wc), ptr(self.cI), ptr(self.vd),
                                                    self.ell.numel(), ptr(self.ell), stream))
        return self.ell[self.sel].contiguous()

    def setup(self, stream):
        A = self.h.lv[0]
        A.vals = self.local_operator(stream)
        self.dist.all_reduce(A.vals)
        self.h.setup_levels()

    def correct(self, r, z):
        """z += E V(E^T (r - A z))  (multiplicative, after the rank-local preconditioner)."""
        e, lib, st, n = self.e, self.lib, self.e._stream(), self.e.ndof
        e.halo.exchange(z)
        capi.check(lib.ech_spmv(n, ptr(e.rowptr), ptr(e.colidx), ptr(e.vals), ptr(z), ptr(self.res), st))
        capi.check(lib.ech_axpby(n, 1.0, ptr(r), -1.0, ptr(self.res), st))             # res = r - A z
        D = self.h.lv[0]
        es = r.element_size()
        rp, ci, v = self.Et
        for a, f in enumerate(self.fields):
            capi.check(lib.ech_spmv(self.nvert, ptr(rp), ptr(ci), ptr(v), capi.C.c_void_p(self.res.data_ptr() + f * self.N * es),
                                    capi.C.c_void_p(D.r.data_ptr() + a * self.nvert * es), st))
        self.dist.all_reduce(D.r)
        self.h.cycle(0, st)
        rp, ci, v = self.E
        for a, f in enumerate(self.fields):
            capi.check(lib.ech_spmv(self.N, ptr(rp), ptr(ci), ptr(v), capi.C.c_void_p(D.z.data_ptr() + a * self.nvert * es),
                                    ptr(self.tmp), st))
            capi.check(lib.ech_axpby(self.N, 1.0, ptr(self.tmp), 1.0, capi.C.c_void_p(z.data_ptr() + f * self.N * es), st))


def _dev_csr(M, dev):
    M = M.tocsr()
    M.sort_indices()
    return (torch.from_numpy(M.indptr.astype(np.int64)).to(dev), torch.from_numpy(M.indices.astype(np.int32)).to(dev),
            torch.from_numpy(M.data.astype(np.float64)).to(dev))


def _torch_csr(M, dev):
    M = M.tocsr()
    M.sort_indices()
    it = np.int32 if max(M.shape) < 2 ** 31 - 1 and M.nnz < 2 ** 31 - 1 else np.int64
    return torch.sparse_csr_tensor(torch.from_numpy(M.indptr.astype(it)).to(dev), torch.from_numpy(M.indices.astype(it)).to(dev),
                                   torch.from_numpy(M.data.astype(np.float64)).to(dev), size=M.shape)
