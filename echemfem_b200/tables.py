"""Reference-element data baked into the generated kernel headers.

Conventions follow what the reference obtains from Firedrake:
quadrature `degree = p + 1` on every measure (echemfem/solver.py:214-251),
i.e. Gauss-Legendre with floor((p+3)/2) points per direction; tensor Lagrange
bases with "spectral" nodes (Gauss-Legendre for DQ, Gauss-Lobatto for Q) or
equispaced nodes behind a switch; lexicographic local numbering, last
direction fastest.

Everything is expressed through ONE 1D table per element: the basis and its
derivative at the q Gauss points followed by the two end points 0 and 1, so
that cell points, own-facet points and neighbour-facet points are all index
lookups into the same (q+2) x (p+1) table.
"""
import numpy as np
from numpy.polynomial import legendre as _leg


def gauss_points(q):
    x, w = _leg.leggauss(q)
    return (x + 1.0) / 2.0, w / 2.0


def lobatto_nodes(m):
    """m Gauss-Lobatto-Legendre nodes on [0,1] by Newton iteration on (1-x^2) P'_{m-1}."""
    if m == 1:
        return np.array([0.5])
    N = m - 1
    x = -np.cos(np.pi * np.arange(m) / N)
    for _ in range(100):
        P = np.zeros((m, m))
        P[:, 0] = 1.0
        P[:, 1] = x
        for k in range(2, m):
            P[:, k] = ((2 * k - 1) * x * P[:, k - 1] - (k - 1) * P[:, k - 2]) / k
        dx = (x * P[:, N] - P[:, N - 1]) / (m * P[:, N])
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    return (x + 1.0) / 2.0


def element_nodes(p, family, variant="spectral"):
    if p == 0:
        return np.array([0.5])
    if variant == "equispaced":
        return np.arange(p + 1) / float(p)
    if variant == "spectral":
        return gauss_points(p + 1)[0] if family == "DG" else lobatto_nodes(p + 1)
    raise ValueError("unknown nodal variant %r" % variant)


def basis_table(nodes, pts):
    """(L, dL)[pt, b] via barycentric weights."""
    nodes = np.asarray(nodes, dtype=float)
    pts = np.asarray(pts, dtype=float)
    m = len(nodes)
    diff = nodes[:, None] - nodes[None, :]
    np.fill_diagonal(diff, 1.0)
    wb = 1.0 / diff.prod(axis=1)
    L = np.zeros((len(pts), m))
    dL = np.zeros((len(pts), m))
    for ip, x in enumerate(pts):
        for b in range(m):
            others = [c for c in range(m) if c != b]
            L[ip, b] = wb[b] * np.prod([x - nodes[c] for c in others])
            s = 0.0
            for c in others:
                s += np.prod([x - nodes[e] for e in others if e != c])
            dL[ip, b] = wb[b] * s
    return L, dL


def num_quad(p):
    return (p + 3) // 2


class ElementTables:
    def __init__(self, p, family, variant="spectral"):
        self.p = p
        self.family = family
        self.variant = variant
        self.nodes = element_nodes(p, family, variant)
        self.q = num_quad(p)
        self.gp, self.gw = gauss_points(self.q)
        self.pts = np.concatenate([self.gp, [0.0, 1.0]])       # index q = face 0, q+1 = face 1
        self.L, self.dL = basis_table(self.nodes, self.pts)
