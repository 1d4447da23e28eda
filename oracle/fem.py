"""Reference-element tables for tensor-product Lagrange elements (oracle).

Restates the external conventions the reference relies on:

* Q1  quadrature: `degree = p + 1` on every measure (echemfem/solver.py:214-251)
      -> Gauss-Legendre with (degree + 2)//2 points per direction.
* Q2  nodal basis: tensor Lagrange; variant "spectral" = Gauss-Legendre nodes
      for discontinuous spaces (DQ), Gauss-Lobatto-Legendre nodes for
      continuous spaces (Q); variant "equispaced" for both.
* Q3  local dof ordering: lexicographic, last index fastest.

Nothing here is used by the product path.
"""
import numpy as np


def gauss_legendre(n):
    """n-point Gauss-Legendre rule on [0, 1]."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto_legendre_nodes(n):
    """n GLL nodes on [0, 1] (n >= 2)."""
    if n == 2:
        return np.array([0.0, 1.0])
    # interior nodes are roots of P'_{n-1}
    c = np.zeros(n)
    c[-1] = 1.0
    dc = np.polynomial.legendre.legder(c)
    r = np.sort(np.polynomial.legendre.legroots(dc))
    x = np.concatenate([[-1.0], r, [1.0]])
    return 0.5 * (x + 1.0)


def nodes_1d(p, family, variant="spectral"):
    """1D nodes of a degree-p Lagrange basis on [0, 1]."""
    if p == 0:
        return np.array([0.5])
    if variant == "equispaced":
        return np.linspace(0.0, 1.0, p + 1)
    if variant != "spectral":
        raise ValueError(variant)
    if family == "DG":
        return gauss_legendre(p + 1)[0]
    return gauss_lobatto_legendre_nodes(p + 1)


def lagrange_1d(nodes, x):
    """Values and derivatives of the Lagrange basis on `nodes` at points x.

    Returns (L, dL) of shape (len(x), len(nodes)).
    """
    x = np.atleast_1d(np.asarray(x, dtype=float))
    m = len(nodes)
    L = np.ones((len(x), m))
    dL = np.zeros((len(x), m))
    for b in range(m):
        den = 1.0
        for c in range(m):
            if c != b:
                den *= nodes[b] - nodes[c]
        for c in range(m):
            if c != b:
                L[:, b] *= x - nodes[c]
        L[:, b] /= den
        for c in range(m):
            if c == b:
                continue
            t = np.ones(len(x))
            for e in range(m):
                if e != b and e != c:
                    t *= x - nodes[e]
            dL[:, b] += t
        dL[:, b] /= den
    return L, dL


def num_quad_points(p):
    """Points per direction for `degree = p + 1` (solver.py:219)."""
    return (p + 1 + 2) // 2


def tabulate(nodes, xi):
    """Tensor basis at reference points xi (npts, d).

    Returns phi (npts, n) and dphi (npts, n, d) with n = len(nodes)**d,
    local index lexicographic, last direction fastest.
    """
    xi = np.asarray(xi, dtype=float)
    npts, d = xi.shape
    m = len(nodes)
    L = []
    dL = []
    for k in range(d):
        a, b = lagrange_1d(nodes, xi[:, k])
        L.append(a)
        dL.append(b)
    n = m ** d
    phi = np.ones((npts, n))
    dphi = np.ones((npts, n, d))
    for b in range(n):
        idx = np.unravel_index(b, (m,) * d)
        for k in range(d):
            phi[:, b] *= L[k][:, idx[k]]
            for l in range(d):
                dphi[:, b, l] *= dL[k][:, idx[k]] if l == k else L[k][:, idx[k]]
    return phi, dphi


def cell_quadrature(p, d):
    """Tensor Gauss-Legendre rule on the reference cell [0,1]^d."""
    x, w = gauss_legendre(num_quad_points(p))
    grids = np.meshgrid(*([x] * d), indexing="ij")
    xi = np.stack([g.ravel() for g in grids], axis=1)
    wg = np.meshgrid(*([w] * d), indexing="ij")
    ww = np.ones_like(wg[0])
    for g in wg:
        ww = ww * g
    return xi, ww.ravel()


def facet_quadrature(p, d):
    """Rule on the reference facet [0,1]^(d-1) (a single point set/weights)."""
    if d == 1:
        return np.zeros((1, 0)), np.ones(1)
    return cell_quadrature(p, d - 1)


def facet_to_cell(lf, t, d):
    """Embed facet-local points t (npts, d-1) of local facet lf into the cell.

    Local facet lf = 2*k + s is the face x_k = s; the remaining coordinates
    keep their order.
    """
    k, s = divmod(lf, 2)
    npts = t.shape[0]
    xi = np.empty((npts, d))
    xi[:, k] = float(s)
    rest = [m for m in range(d) if m != k]
    for a, m in enumerate(rest):
        xi[:, m] = t[:, a]
    return xi
