"""scipy prototype of the multigrid cycle of echemfem_b200/multigrid.py + auxspace.py on ORACLE Jacobians (CPU only).

Test infrastructure: this is the analysis the preconditioner's design rests on (the numbers quoted in
echemfem_b200/auxspace.py and DESIGN.md 3.4b come from it), kept runnable so that the claims stay checkable without
a GPU.  Problem: the cfg2 physics (electroneutral Nernst-Planck MMS, D = (0.5e-5, 1e-5), fields (C1, Phi)) on
N x N quads, DG-Q1; operators from oracle/assemble.py.  Cycle: nested DG-Q1 hierarchy with Galerkin operators,
block-Jacobi / ILU(0) smoothers on tiles of cells, V(1,1), exact coarsest solve; optionally the continuous-Q1
auxiliary-space correction for chosen fields, added at the finest level from the same residual; right-preconditioned
GMRES to 1e-12 on a random right-hand side.

    python tests/mg_prototype.py 64            # iteration counts of the variants at N = 64
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

_ILU_C = r"""
#include <stdlib.h>
/* in-place ILU(0) of a CSR matrix with sorted columns; diag[i] = position of the diagonal entry */
int ilu0(int n, const long *rp, const int *ci, double *v, long *diag) {
    int *pos = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) pos[i] = -1;
    for (int i = 0; i < n; ++i) {
        diag[i] = -1;
        for (long p = rp[i]; p < rp[i + 1]; ++p) { pos[ci[p]] = (int)(p - rp[i]); if (ci[p] == i) diag[i] = p; }
        if (diag[i] < 0) { free(pos); return 1; }
        for (long p = rp[i]; p < rp[i + 1] && ci[p] < i; ++p) {
            int k = ci[p];
            double l = v[p] / v[diag[k]];
            v[p] = l;
            for (long q = diag[k] + 1; q < rp[k + 1]; ++q) {
                int j = ci[q];
                if (pos[j] >= 0) v[rp[i] + pos[j]] -= l * v[q];
            }
        }
        for (long p = rp[i]; p < rp[i + 1]; ++p) pos[ci[p]] = -1;
    }
    free(pos);
    return 0;
}
void ilu0_solve(int n, const long *rp, const int *ci, const double *v, const long *diag, const double *r, double *z) {
    for (int i = 0; i < n; ++i) {
        double s = r[i];
        for (long p = rp[i]; p < diag[i]; ++p) s -= v[p] * z[ci[p]];
        z[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = z[i];
        for (long p = diag[i] + 1; p < rp[i + 1]; ++p) s -= v[p] * z[ci[p]];
        z[i] = s / v[diag[i]];
    }
}
"""
_lib = None


def _ilu_lib():
    global _lib
    if _lib is None:
        d = os.path.join(ROOT, "oracle", "_build")
        os.makedirs(d, exist_ok=True)
        src, so = os.path.join(d, "mg_prototype_ilu0.c"), os.path.join(d, "libmg_prototype_ilu0.so")
        if not os.path.exists(so) or not os.path.exists(src) or open(src).read() != _ILU_C:
            with open(src, "w") as f:
                f.write(_ILU_C)
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", src, "-o", so])
        _lib = C.CDLL(so)
        _lib.ilu0.argtypes = [C.c_int] + [C.c_void_p] * 4
        _lib.ilu0_solve.argtypes = [C.c_int] + [C.c_void_p] * 6
    return _lib


D1, D2 = 0.5e-5, 1.0e-5


def make_disc(N, R=1):
    """cfg2 physics on [0, R] x [0, 1] with (R N) x N square cells (R = 1: the unit square of the benchmark; R > 1: the
    domain of the weak-scaling runs, R strips side by side along the flow)."""
    import mms_cases
    from oracle.mesh import tensor_mesh
    from oracle.assemble import Discretization
    d = Discretization(tensor_mesh([np.linspace(0.0, float(R), R * N + 1), np.linspace(0.0, 1.0, N + 1)]),
                       mms_cases.adm_problem(1.0, D1=D1, D2=D2))
    d.R = R
    return d


def state(disc):
    import mms_cases
    x = disc.node_coords
    pert = 1e-3 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    return np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.Uex(x) - pert])


def cell_ij(disc, N):
    xc = disc.Xc.mean(axis=1)
    return np.minimum((xc * N).astype(int), np.array([getattr(disc, "R", 1) * N - 1, N - 1]))


def dg_prolongation(fd, Nf, cd, Nc):
    """Scalar embedding of the coarse DG-Q1 space (2 x 2 merged cells): coarse bilinear basis at the fine nodes."""
    fij, cij = cell_ij(fd, Nf), cell_ij(cd, Nc)
    cmap = -np.ones((getattr(cd, "R", 1) * Nc, Nc), dtype=int)
    cmap[cij[:, 0], cij[:, 1]] = np.arange(len(cij))
    parent = cmap[fij[:, 0] // 2, fij[:, 1] // 2]
    xf = fd.node_coords.reshape(-1, 4, 2)
    xcn = cd.node_coords.reshape(-1, 4, 2)
    # the coarse nodal basis is Lagrange on the coarse cell's own (Gauss) nodes: evaluate through its 1D factors
    rows, cols, vals = [], [], []
    for a in range(4):
        for b in range(4):
            w = np.ones(len(fij))
            for k in range(2):
                xb = xcn[parent, b, k]
                others = [xcn[parent, bb, k] for bb in range(4) if abs(bb - b) == (2 if k == 0 else 1) and (bb // 2 == b // 2 if k == 1 else bb % 2 == b % 2)]
                xo = others[0]
                w *= (xf[:, a, k] - xo) / (xb - xo)
            rows.append(np.arange(len(fij)) * 4 + a)
            cols.append(parent * 4 + b)
            vals.append(w)
    P = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(fd.ndof_scalar, cd.ndof_scalar))
    P.eliminate_zeros()
    return P


def cg_embedding(disc, N):
    """Q: continuous Q1 vertex values -> values at the DG nodes ((N+1)^2 vertices, lexicographic)."""
    x, h = disc.node_coords, 1.0 / N
    ij = np.repeat(cell_ij(disc, N), 4, axis=0)
    xi = x / h - ij
    rows, cols, vals = [], [], []
    for vx in (0, 1):
        for vy in (0, 1):
            rows.append(np.arange(len(x)))
            cols.append((ij[:, 0] + vx) * (N + 1) + ij[:, 1] + vy)
            vals.append((xi[:, 0] if vx else 1 - xi[:, 0]) * (xi[:, 1] if vy else 1 - xi[:, 1]))
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(len(x), (getattr(disc, "R", 1) * N + 1) * (N + 1)))


def tiles(disc, N, tx, ty):
    """Dof index sets of the tx x ty cell tiles (both fields, field-major inside the tile)."""
    ij = cell_ij(disc, N)
    tid = (ij[:, 0] // tx) * (N // ty + 1) + ij[:, 1] // ty
    ns = disc.ndof_scalar
    out = []
    for t in np.unique(tid):
        cells = np.nonzero(tid == t)[0]
        cells = cells[np.lexsort((ij[cells, 1], ij[cells, 0]))]
        d0 = (cells[:, None] * 4 + np.arange(4)[None, :]).ravel()
        out.append(np.concatenate([d0, d0 + ns]))
    return out


class TileILU:
    """Block Jacobi over index sets, ILU(0) inside each block (one permuted block-diagonal matrix)."""

    def __init__(self, A, tl):
        A = A.tocsr()
        self.perm = np.concatenate(tl)
        n = A.shape[0]
        blk = np.empty(n, dtype=np.int64)
        for t, idx in enumerate(tl):
            blk[idx] = t
        Ap = A[self.perm][:, self.perm].tocoo()
        bp = blk[self.perm]
        keep = bp[Ap.row] == bp[Ap.col]
        B = sp.csr_matrix((Ap.data[keep], (Ap.row[keep], Ap.col[keep])), shape=A.shape)
        B.sort_indices()
        self.rp, self.ci, self.v = B.indptr.astype(np.int64), B.indices.astype(np.int32), B.data.copy()
        self.diag = np.zeros(n, dtype=np.int64)
        self.n = n
        assert _ilu_lib().ilu0(n, self.rp.ctypes.data, self.ci.ctypes.data, self.v.ctypes.data, self.diag.ctypes.data) == 0

    def solve(self, r):
        rp_ = np.ascontiguousarray(r[self.perm])
        zp = np.empty_like(rp_)
        _ilu_lib().ilu0_solve(self.n, self.rp.ctypes.data, self.ci.ctypes.data, self.v.ctypes.data, self.diag.ctypes.data,
                              rp_.ctypes.data, zp.ctypes.data)
        z = np.empty_like(zp)
        z[self.perm] = zp
        return z


def interp1d(Nf):
    Nc = Nf // 2
    P = sp.lil_matrix((Nf + 1, Nc + 1))
    for i in range(Nf + 1):
        if i % 2 == 0:
            P[i, i // 2] = 1.0
        else:
            P[i, i // 2] = 0.5
            P[i, i // 2 + 1] = 0.5
    return P.tocsr()


class ConformingMG:
    """Bilinear interpolation, Galerkin operators, damped Jacobi V(1,1), exact coarsest level."""

    def __init__(self, A, N, nfields=1, omega=0.8, coarsest=8, R=1):
        self.A, self.P, self.omega = [A.tocsr()], [], omega
        while N > coarsest and N % 2 == 0:
            P = sp.kron(interp1d(R * N), interp1d(N), format="csr")
            P = sp.block_diag([P] * nfields, format="csr") if nfields > 1 else P
            self.P.append(P)
            self.A.append((P.T @ self.A[-1] @ P).tocsr())
            N //= 2
        self.lu = spla.splu(self.A[-1].tocsc())
        self.Dinv = [1.0 / M.diagonal() for M in self.A]

    def solve(self, b, l=0):
        if l == len(self.A) - 1:
            return self.lu.solve(b)
        x = self.omega * self.Dinv[l] * b
        x = x + self.P[l] @ self.solve(self.P[l].T @ (b - self.A[l] @ x), l + 1)
        return x + self.omega * self.Dinv[l] * (b - self.A[l] @ x)


class Cycle:
    def __init__(self, N, tile=(8, 8), aux_fields=(), coarsest=8, R=1):
        discs, Ns = [make_disc(N, R)], [N]
        while Ns[-1] > coarsest:
            Ns.append(Ns[-1] // 2)
            discs.append(make_disc(Ns[-1], R))
        self.R = R
        self.discs, self.Ns = discs, Ns
        self.A = [discs[0].jacobian(state(discs[0])).tocsr()]
        self.P = []
        for l in range(len(Ns) - 1):
            P = sp.block_diag([dg_prolongation(discs[l], Ns[l], discs[l + 1], Ns[l + 1])] * 2, format="csr")
            self.P.append(P)
            self.A.append((P.T @ self.A[-1] @ P).tocsr())
        self.S = [TileILU(self.A[l], tiles(discs[l], Ns[l], min(tile[0], Ns[l]), min(tile[1], Ns[l])))
                  for l in range(len(Ns) - 1)]
        self.coarse = spla.splu(self.A[-1].tocsc())
        self.aux = None
        if aux_fields:
            Q = cg_embedding(discs[0], N)
            Z = sp.csr_matrix(Q.shape)
            self.Pa = sp.vstack([sp.hstack([(Q if f == g else Z) for g in aux_fields]) for f in range(2)]).tocsr()
            self.aux = ConformingMG(self.Pa.T @ self.A[0] @ self.Pa, N, nfields=len(aux_fields), R=R)

    def vcycle(self, b, l=0):
        if l == len(self.Ns) - 1:
            return self.coarse.solve(b)
        A, S = self.A[l], self.S[l]
        x = S.solve(b)
        r = b - A @ x
        x = x + self.P[l] @ self.vcycle(self.P[l].T @ r, l + 1)
        if l == 0 and self.aux is not None:
            x = x + self.Pa @ self.aux.solve(self.Pa.T @ r)
        return x + S.solve(b - A @ x)


def gmres_iterations(A, b, M, rtol=1e-12, maxit=300):
    """Iterations of full right-preconditioned GMRES (Gram-Schmidt with one refinement pass) to rtol; maxit if not reached."""
    beta = np.linalg.norm(b)
    V, H = [b / beta], np.zeros((maxit + 1, maxit))
    for k in range(maxit):
        w = A @ M(V[k])
        for _ in range(2):
            for i in range(k + 1):
                h = w @ V[i]
                H[i, k] += h
                w = w - h * V[i]
        H[k + 1, k] = np.linalg.norm(w)
        V.append(w / H[k + 1, k])
        e1 = np.zeros(k + 2)
        e1[0] = beta
        y = np.linalg.lstsq(H[:k + 2, :k + 1], e1, rcond=None)[0]
        if np.linalg.norm(H[:k + 2, :k + 1] @ y - e1) < rtol * beta:
            return k + 1
    return maxit


def block_iterations(N, field, tile=(8, 8), maxit=300):
    """The same cycle for ONE diagonal block of the system (field 0: concentration, 1: potential)."""
    c = Cycle(N, tile=tile)
    sub = Cycle.__new__(Cycle)
    sub.Ns, sub.discs, sub.aux = c.Ns, c.discs, None
    sub.A, sub.P, sub.S = [], [], []
    for l, A in enumerate(c.A):
        n = c.discs[l].ndof_scalar
        sl = slice(field * n, (field + 1) * n)
        sub.A.append(A[sl][:, sl].tocsr())
    for l, P in enumerate(c.P):
        n, nc = c.discs[l].ndof_scalar, c.discs[l + 1].ndof_scalar
        sub.P.append(P[:n][:, :nc].tocsr())
    for l in range(len(c.Ns) - 1):
        tl = tiles(c.discs[l], c.Ns[l], min(tile[0], c.Ns[l]), min(tile[1], c.Ns[l]))
        sub.S.append(TileILU(sub.A[l], [t[:len(t) // 2] for t in tl]))
    sub.coarse = spla.splu(sub.A[-1].tocsc())
    b = np.random.default_rng(0).standard_normal(sub.A[0].shape[0])
    return gmres_iterations(sub.A[0], b, sub.vcycle, maxit=maxit)


def rediscretised_iterations(N, tile=(8, 8), maxit=300):
    """Coarse operators assembled on the coarse meshes (at the interpolated state) instead of Galerkin products."""
    c = Cycle(N, tile=tile)
    for l in range(1, len(c.Ns)):
        c.A[l] = c.discs[l].jacobian(state(c.discs[l])).tocsr()
    c.S = [TileILU(c.A[l], tiles(c.discs[l], c.Ns[l], min(tile[0], c.Ns[l]), min(tile[1], c.Ns[l])))
           for l in range(len(c.Ns) - 1)]
    c.coarse = spla.splu(c.A[-1].tocsc())
    b = np.random.default_rng(0).standard_normal(c.A[0].shape[0])
    return gmres_iterations(c.A[0], b, c.vcycle, maxit=maxit)


def two_rank_iterations(N, maxit=300, ranks=2, R=1):
    """The mesh cut into `ranks` strips along x, as the ranks would hold it: (one rank, block Jacobi of the
    rank-local cycles -- couplings across the cuts dropped, local conforming spaces included --, the same plus the
    conforming space on the GLOBAL vertex grid applied multiplicatively afterwards).  R: domain [0, R] x [0, 1]
    (R = ranks: the weak-scaling layout, one N x N strip per rank)."""
    c = Cycle(N, aux_fields=(1,), R=R)
    A = c.A[0]
    rank_of_dof = np.tile(np.repeat(cell_ij(c.discs[0], N)[:, 0] // (R * N // ranks), 4), 2)
    Ac = A.tocoo()
    keep = rank_of_dof[Ac.row] == rank_of_dof[Ac.col]
    Aloc = sp.csr_matrix((Ac.data[keep], (Ac.row[keep], Ac.col[keep])), shape=A.shape)
    loc = Cycle.__new__(Cycle)
    loc.Ns, loc.discs, loc.P, loc.Pa = c.Ns, c.discs, c.P, c.Pa
    loc.A = [Aloc]
    for P in c.P:
        loc.A.append((P.T @ loc.A[-1] @ P).tocsr())
    loc.S = [TileILU(loc.A[l], tiles(c.discs[l], c.Ns[l], min(8, c.Ns[l]), min(8, c.Ns[l]))) for l in range(len(c.Ns) - 1)]
    loc.coarse = spla.splu(loc.A[-1].tocsc())
    loc.aux = ConformingMG(c.Pa.T @ Aloc @ c.Pa, N, R=R)
    G = ConformingMG(c.Pa.T @ A @ c.Pa, N, R=R)

    def with_global(r):
        z = loc.vcycle(r)
        return z + c.Pa @ G.solve(c.Pa.T @ (r - A @ z))

    b = np.random.default_rng(0).standard_normal(A.shape[0])
    return (gmres_iterations(A, b, c.vcycle, maxit=maxit), gmres_iterations(A, b, loc.vcycle, maxit=maxit),
            gmres_iterations(A, b, with_global, maxit=maxit))


def iterations(N, aux_fields=(), tile=(8, 8), maxit=300):
    c = Cycle(N, tile=tile, aux_fields=aux_fields)
    b = np.random.default_rng(0).standard_normal(c.A[0].shape[0])
    return gmres_iterations(c.A[0], b, c.vcycle, maxit=maxit)


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    print("N = %d: nested DG hierarchy alone %d iterations; + conforming space for the potential %d; "
          "+ conforming space for both fields %d (cap 300)" % (
              N, iterations(N), iterations(N, (1,)), iterations(N, (0, 1))))
    print("        concentration block alone %d, potential block alone %d; rediscretised coarse levels %d; "
          "tiles 32 x 2 / 2 x 32 (with the conforming space): %d / %d" % (
              block_iterations(N, 0), block_iterations(N, 1), rediscretised_iterations(N),
              iterations(N, (1,), tile=(32, 2)), iterations(N, (1,), tile=(2, 32))))
