"""Mesh inputs of the benchmark configurations that the reference obtains from gmsh.

`examples/bortels_structuredquad_nondim.geo` of the reference is a
transfinite (structured) quadrilateral mesh of the channel [0, 12] x [0, 1] whose curves carry gmsh's
"Bump" grading.  gmsh is not available offline, so the grading law is restated here:

    Transfinite Curve{...} = n Using Bump r    places n nodes such that the local spacing along the
    curve of length L is   h(s) = b - a (s - L/2)^2,  with
        a = 2 sqrt(1-r) log|(1 + 1/sqrt(1-r)) / (1 - 1/sqrt(1-r))| / (n L)          (r < 1)
        a = -4 sqrt(r-1) atan2(1, sqrt(r-1)) / (n L)                                 (r > 1)
        b = -a L^2 / (4 (r - 1)),
    nodes sitting at equal increments of  int ds / h(s).

Node positions therefore agree with gmsh's up to its quadrature of that integral (parity of the mesh
itself is unpinned; the physics and boundary ids are exact).
"""
import numpy as np

from .mesh import TensorMesh


def bump_nodes(length, n, coef, samples=200001):
    """n node abscissae on [0, length] with gmsh's `Using Bump coef` distribution."""
    if coef <= 0.0 or coef == 1.0:
        return np.linspace(0.0, length, n)
    if coef > 1.0:
        a = -4.0 * np.sqrt(coef - 1.0) * np.arctan2(1.0, np.sqrt(coef - 1.0)) / (n * length)
    else:
        q = np.sqrt(1.0 - coef)
        a = 2.0 * q * np.log(abs((1.0 + 1.0 / q) / (1.0 - 1.0 / q))) / (n * length)
    b = -a * length * length / (4.0 * (coef - 1.0))
    s = np.linspace(0.0, length, samples)
    dens = 1.0 / (-a * (s - 0.5 * length) ** 2 + b)
    cum = np.concatenate([[0.0], np.cumsum(0.5 * (dens[1:] + dens[:-1]) * np.diff(s))])
    targets = np.linspace(0.0, cum[-1], n)
    x = np.interp(targets, cum, s)
    x[0], x[-1] = 0.0, length
    return x


def bortels_axes(n_in=100, n_el=100, n_out=100, n_y=50, La=5.0, L=2.0, Lb=5.0, h=1.0):
    """Vertex abscissae of bortels_structuredquad_nondim.geo (lines 33-36: 100/100/100 x 50 nodes)."""
    xa = bump_nodes(La, n_in, 0.03)
    xe = La + bump_nodes(L, n_el, 0.1)
    xb = La + L + bump_nodes(Lb, n_out, 0.03)
    x = np.concatenate([xa, xe[1:], xb[1:]])
    y = bump_nodes(h, n_y, 0.03)
    return x, y


def bortels_markers(La=5.0, L=2.0, Lb=5.0, h=1.0):
    """Physical curves of the .geo: Inlet 10 (x=0), Outlet 11 (x=La+L+Lb), Anode 12 (top, electrode
    segment), Cathode 13 (bottom, electrode segment), Wall 14 (the rest)."""
    def fn(fcoords, fverts):
        xm = fcoords[:, :, 0].mean(axis=1)
        ym = fcoords[:, :, 1].mean(axis=1)
        dx = np.abs(fcoords[:, 0, 0] - fcoords[:, 1, 0])
        vertical = dx < 1e-14
        tol = 1e-12
        out = np.full(fcoords.shape[0], 14, dtype=np.int64)
        out[vertical & (np.abs(xm) < tol)] = 10
        out[vertical & (np.abs(xm - (La + L + Lb)) < tol)] = 11
        electrode = (~vertical) & (xm > La) & (xm < La + L)
        out[electrode & (np.abs(ym - h) < tol)] = 12
        out[electrode & (np.abs(ym) < tol)] = 13
        return out
    return fn


def BortelsMesh(n_in=100, n_el=100, n_out=100, n_y=50):
    """297 x 49 = 14 553 cells with the default node counts."""
    x, y = bortels_axes(n_in, n_el, n_out, n_y)
    return TensorMesh([x, y], name="bortels_structuredquad_nondim", marker_of_facet=bortels_markers())


# ---------------------------------------------------------------------------------------------------
# Boundary-layer meshes (echemfem/utility_meshes.py:9-195): a uniform core plus uniformly refined layers
# next to the marked boundaries.  The node abscissae follow the reference's arithmetic (same linspace /
# arange end points), so node positions are bit-identical; the cells are numbered by TensorMesh.

def _layered_axis(n, length, n_layer, l_layer, lo_layer, hi_layer):
    """n uniform cells between the layers, n_layer cells of total width l_layer in each requested layer."""
    a0, a1 = 0.0, float(length)
    if lo_layer:
        a0 += l_layer
    if hi_layer:
        a1 -= l_layer
    ax = np.linspace(a0, a1, n + 1, dtype=np.double)
    if lo_layer:
        ax = np.append(np.linspace(0.0, l_layer - l_layer / n_layer, n_layer, dtype=np.double), ax)
    if hi_layer:
        ax = np.append(ax, np.linspace(a1 + l_layer / n_layer, length, n_layer, dtype=np.double))
    return ax


def RectangleBoundaryLayerMesh(nx, ny, Lx, Ly, n_bdlayer, L_bdlayer, Ly_bdlayer=None, ny_bdlayer=None,
                               boundary=(1,), reorder=None, distribution_parameters=None, comm=None, tile=None):
    """Quadrilateral mesh of [0, Lx] x [0, Ly] with boundary layers next to the sides listed in `boundary`
    (1: x = 0, 2: x = Lx, 3: y = 0, 4: y = Ly); same signature as echemfem/utility_meshes.py:75-78."""
    from .mesh import TensorMesh
    for n in (nx, ny):
        if n <= 0 or n % 1:
            raise ValueError("Number of cells must be a postive integer")
    lx_layer = L_bdlayer
    ly_layer = L_bdlayer if Ly_bdlayer is None else Ly_bdlayer
    nx_layer = n_bdlayer
    ny_layer = n_bdlayer if ny_bdlayer is None else ny_bdlayer
    xs = _layered_axis(nx, Lx, nx_layer, lx_layer, 1 in boundary, 2 in boundary)
    ys = _layered_axis(ny, Ly, ny_layer, ly_layer, 3 in boundary, 4 in boundary)
    return TensorMesh([xs, ys], name="rectangle_boundary_layer", tile=tile)


def IntervalBoundaryLayerMesh(ncells, length_or_left, ncells_bdlayer, length_bdlayer, boundary=(2,), right=None,
                              distribution_parameters=None, comm=None):
    """Interval mesh with boundary layers at the ends listed in `boundary` (1: left, 2: right);
    same signature as echemfem/utility_meshes.py:9-10."""
    from .mesh import TensorMesh
    if right is None:
        left, right = 0.0, float(length_or_left)
    else:
        left = float(length_or_left)
    if ncells <= 0 or ncells % 1:
        raise ValueError("Number of cells must be a postive integer")
    a0, a1 = left, right
    if 1 in boundary:
        a0 += length_bdlayer
    if 2 in boundary:
        a1 -= length_bdlayer
    if a1 - a0 < 0:
        raise ValueError("Requested mesh has negative length")
    h = (a1 - a0) / ncells
    h1 = length_bdlayer / ncells_bdlayer
    ax = np.arange(a0, a1 + 0.01 * h, h, dtype=np.double)
    if 1 in boundary:
        ax = np.append(np.arange(left, length_bdlayer - 0.99 * h1, h1, dtype=np.double), ax)
    if 2 in boundary:
        ax = np.append(ax, np.arange(right - length_bdlayer + h1, right + 0.01 * h1, h1, dtype=np.double))
    return TensorMesh([ax], name="interval_boundary_layer")
