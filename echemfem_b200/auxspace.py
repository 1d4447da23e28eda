"""Continuous-Q1 auxiliary space for the potential fields of the DG multigrid preconditioner.

Why.  The Galerkin coarse operators of nested DG spaces inherit the interior-penalty term of the FINE mesh
(sigma / h_fine on every coarse face), so on level l the jump penalty is 2^l times too strong for that level and the
smoothers only see the jump modes: the V-cycle of `multigrid.py` loses a constant factor per level on the elliptic
(potential) block -- measured on the cfg2 problem: GMRES iterations 42 / 75 / 142 at N = 32 / 64 / 128, the potential
block alone 70 / 118 at N = 64 / 128, the (advective) concentration block alone 12 / 22.  A CONTINUOUS coarse function
has no jumps, so the Galerkin operator of the conforming subspace carries no penalty at all and its multigrid
hierarchy (bilinear interpolation, damped Jacobi) is the textbook one.  The correction
        z_phi += Q V_cg(Q^T r_phi),      Q: continuous Q1 vertex values -> values at the DG nodes (the embedding),
is added at the finest level next to the DG coarse-grid correction (same residual: no extra fine-level SpMV).  With it
a scipy prototype of the same cycle needs 19 / 24 iterations at N = 64 / 128 (73 / 142 without).  Only the potential fields
get it: for the advection-dominated concentration fields the conforming Galerkin operator is the central (unstable)
discretisation and the correction diverges (measured), their nested DG hierarchy keeps the upwinding.

The reference reaches the same kind of hierarchy through PETSc (`gmg`: geometric multigrid on a MeshHierarchy, `amg`:
hypre on the assembled operator, echemfem/solver.py:1190-1213); this is the hand-built equivalent for tensor meshes.

Everything in the apply path is this package's own kernels (`ech_spmv`, `ech_axpby`, `ech_dense_gemv`); the set-up
uses `ech_aux_galerkin` for Q^T A Q (cell block by cell block, fixed pattern) and cuSPARSE SpGEMM through `torch.sparse.mm`
for the small conforming coarse levels (<= 1/16 of the fine matrix in total).
"""
import numpy as np
import scipy.sparse as sp
import torch

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


class AuxSpaceCG:
    def __init__(self, mg, fields, omega=0.8, coarsest_nodes=400):
        e = mg.e
        self.mg, self.lib, self.dev = mg, mg.lib, mg.dev
        self.fields = [int(f) for f in fields]
        self.na = len(self.fields)
        self.omega = float(omega)
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
        # conforming hierarchy: vertex sets coarsened like the cells of multigrid.Hierarchy
        self.lv = []
        while True:
            D = _Lv()
            D.nnode = int(np.prod([len(a) for a in axes]))
            D.n = D.nnode * self.na
            for name in ("r", "z", "t", "s"):
                setattr(D, name, torch.zeros(D.n, dtype=torch.float64, device=self.dev))
            self.lv.append(D)
            if D.nnode <= coarsest_nodes or any(len(a) < 4 for a in axes) or len(self.lv) >= 14:
                break
            cax = [a[::2] if (len(a) - 1) % 2 == 0 else np.append(a[::2], a[-1]) for a in axes]
            P = interpolation_1d(axes[0], cax[0])
            for k in range(1, len(axes)):
                P = sp.kron(P, interpolation_1d(axes[k], cax[k]), format="csr")
            D.P = _dev_csr(P, self.dev)
            D.R = _dev_csr(P.T, self.dev)
            Pb = sp.block_diag([P] * self.na, format="csr") if self.na > 1 else P.tocsr()
            D.Pt = _torch_csr(Pb, self.dev)
            D.Rt = _torch_csr(Pb.T.tocsr(), self.dev)
            D.ncoarse = int(np.prod([len(a) for a in cax]))
            axes = cax
        self._pattern = None

    # -- set-up --------------------------------------------------------------------------------------------------
    def _symbolic(self, rowptr, colidx):
        """Cell blocks of the fine matrix that couple two auxiliary fields on owned cells, and where their
        W^T B W lands in A_aux.  A block is found by its first entry (row node 0, column node 0)."""
        N, ncg, na, nv = self.N, self.ncg, self.na, self.nv
        nrow = na * ncg
        vt = self.vtab.view(-1, nv)
        row0_all, off_all, key_all = [], [], []
        chunk_rows = 1 << 20
        for ia, fa in enumerate(self.fields):
            for r0 in range(fa * N, (fa + 1) * N, chunk_rows):
                r1 = min((fa + 1) * N, r0 + chunk_rows)
                p0, p1 = int(rowptr[r0].item()), int(rowptr[r1].item())
                if p1 == p0:
                    continue
                counts = rowptr[r0 + 1:r1 + 1] - rowptr[r0:r1]
                rows = torch.repeat_interleave(torch.arange(r0, r1, device=self.dev), counts)
                cols = colidx[p0:p1].to(torch.int64)
                cf = cols // N
                ja = torch.full_like(cf, -1)
                for k, fb in enumerate(self.fields):
                    ja[cf == fb] = k
                rl, cl = rows - fa * N, cols - cf * N
                head = (ja >= 0) & (rl % nv == 0) & (cl % nv == 0)
                pos = torch.nonzero(head).reshape(-1)
                rl, cl, jsel, rsel = rl[pos], cl[pos], ja[pos], rows[pos]
                vi, vj = vt[rl // nv], vt[cl // nv]                         # (heads, nv) vertex numbers
                ok = (vi[:, 0] >= 0) & (vj[:, 0] >= 0)                        # owned cells on both sides
                pos, vi, vj, jsel, rsel = pos[ok], vi[ok], vj[ok], jsel[ok], rsel[ok]
                row0_all.append(rsel)
                off_all.append((pos + p0 - rowptr[rsel]).to(torch.int32))
                key_all.append(((ia * ncg + vi)[:, :, None] * nrow + (jsel * ncg)[:, None, None] + vj[:, None, :]).reshape(-1))
                del rows, cols, cf, ja, rl, cl, head, pos, vi, vj, ok, jsel, rsel
        keys = torch.cat(key_all)
        del key_all
        uniq, inv = torch.unique(keys, return_inverse=True)
        del keys
        A = self.lv[0]
        A.colidx = (uniq % nrow).to(torch.int32).contiguous()
        rcount = torch.bincount(uniq // nrow, minlength=nrow)
        A.rowptr = torch.zeros(nrow + 1, dtype=torch.int64, device=self.dev)
        A.rowptr[1:] = torch.cumsum(rcount, 0)
        A.vals = torch.zeros(uniq.numel(), dtype=torch.float64, device=self.dev)
        self.row0 = torch.cat(row0_all).contiguous()
        self.off = torch.cat(off_all).contiguous()
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
        for l in range(len(self.lv) - 1):
            D, Dc = self.lv[l], self.lv[l + 1]
            idt = D.Pt.crow_indices().dtype
            At = torch.sparse_csr_tensor(D.rowptr.to(idt), D.colidx.to(idt), D.vals, size=(D.n, D.n))
            Ac = torch.sparse.mm(D.Rt, torch.sparse.mm(At, D.Pt))
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
            D.w = (self.omega / diag).contiguous()
            if not hasattr(D, "wrow") or D.wrow.numel() != D.n + 1:
                D.wrow = torch.arange(D.n + 1, dtype=torch.int64, device=self.dev)
                D.wcol = torch.arange(D.n, dtype=torch.int32, device=self.dev)
        C = self.lv[-1]
        dense = torch.sparse_csr_tensor(C.rowptr, C.colidx.to(torch.int64), C.vals, size=(C.n, C.n)).to_dense()
        C.inv = torch.linalg.inv(dense).contiguous()

    # -- apply ---------------------------------------------------------------------------------------------------
    def _spmv(self, rp, ci, v, n, x, y, st):
        capi.check(self.lib.ech_spmv(n, ptr(rp), ptr(ci), ptr(v), ptr(x), ptr(y), st))

    def _transfer(self, M, nrow, x, xstride, y, ystride, st):
        rp, ci, v = M
        es = x.element_size()
        for a in range(self.na):
            capi.check(self.lib.ech_spmv(nrow, ptr(rp), ptr(ci), ptr(v), capi.C.c_void_p(x.data_ptr() + a * xstride * es),
                                         capi.C.c_void_p(y.data_ptr() + a * ystride * es), st))

    def _cycle(self, l, st):
        lib = self.lib
        D = self.lv[l]
        if l == len(self.lv) - 1:
            capi.check(lib.ech_dense_gemv(D.n, D.n, ptr(D.inv), ptr(D.r), ptr(D.z), st))
            return
        Dc = self.lv[l + 1]
        self._spmv(D.wrow, D.wcol, D.w, D.n, D.r, D.z, st)                       # z = omega D^-1 r
        self._spmv(D.rowptr, D.colidx, D.vals, D.n, D.z, D.t, st)
        capi.check(lib.ech_axpby(D.n, 1.0, ptr(D.r), -1.0, ptr(D.t), st))         # t = r - A z
        self._transfer(D.R, Dc.nnode, D.t, D.nnode, Dc.r, Dc.nnode, st)
        self._cycle(l + 1, st)
        self._transfer(D.P, D.nnode, Dc.z, Dc.nnode, D.t, D.nnode, st)
        capi.check(lib.ech_axpby(D.n, 1.0, ptr(D.t), 1.0, ptr(D.z), st))          # z += P z_c
        self._spmv(D.rowptr, D.colidx, D.vals, D.n, D.z, D.t, st)
        capi.check(lib.ech_axpby(D.n, 1.0, ptr(D.r), -1.0, ptr(D.t), st))         # t = r - A z
        self._spmv(D.wrow, D.wcol, D.w, D.n, D.t, D.s, st)
        capi.check(lib.ech_axpby(D.n, 1.0, ptr(D.s), 1.0, ptr(D.z), st))          # z += omega D^-1 t

    def restrict_and_solve(self, res, st):
        """One V-cycle of the auxiliary problem for the fine-level residual `res` (all fields, stride N)."""
        D = self.lv[0]
        es = res.element_size()
        rp, ci, v = self.Qt
        for a, f in enumerate(self.fields):
            capi.check(self.lib.ech_spmv(self.ncg, ptr(rp), ptr(ci), ptr(v), capi.C.c_void_p(res.data_ptr() + f * self.N * es),
                                         capi.C.c_void_p(D.r.data_ptr() + a * self.ncg * es), st))
        self._cycle(0, st)

    def prolong_add(self, z, tmp, st):
        """z_f += Q z_aux for every auxiliary field f; tmp: scratch of at least N entries."""
        D = self.lv[0]
        es = z.element_size()
        rp, ci, v = self.Q
        for a, f in enumerate(self.fields):
            capi.check(self.lib.ech_spmv(self.N, ptr(rp), ptr(ci), ptr(v), capi.C.c_void_p(D.z.data_ptr() + a * self.ncg * es),
                                         ptr(tmp), st))
            capi.check(self.lib.ech_axpby(self.N, 1.0, ptr(tmp), 1.0, capi.C.c_void_p(z.data_ptr() + f * self.N * es), st))


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
