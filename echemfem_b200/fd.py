"""The Firedrake names the reference's scripts use (`from firedrake import *`).

SURVEY.md R1: example and test scripts of the reference start with
`from firedrake import *` and build their physics closures from ~35 of its
names.  This module provides those names on top of `echemfem_b200.symbolic`,
`.mesh` and `.function`, so that such a script runs against this package after
`import echemfem_b200.compat` (which installs it as `firedrake` when the real
one is absent).
"""
from .symbolic import (Constant, grad, div, inner, dot, outer, avg, jump, as_vector, conditional,  # noqa: F401
                       exp, ln, sqrt, sin, cos, tan, tanh, sinh, cosh, sign, max_value, min_value,
                       gt, lt, ge, le, And, Or, Not, pi, dx, ds, dS, Measure)
from .symbolic import coordinate as _coordinate, facet_normal as _facet_normal
from .symbolic import cell_volume as _cv, cell_diameter as _cd, facet_area as _fa
from .mesh import (Mesh as _Mesh, TensorMesh, UnitSquareMesh, RectangleMesh, UnitCubeMesh,  # noqa: F401
                   ExtrudedMesh)
from .function import (Function, FunctionSpace, MixedFunctionSpace, errornorm)  # noqa: F401


def SpatialCoordinate(mesh):
    return _coordinate(mesh.dim)


def FacetNormal(mesh):
    return _facet_normal(mesh.dim)


def CellVolume(mesh):
    return _cv()


def CellDiameter(mesh):
    return _cd()


def FacetArea(mesh):
    return _fa()


def Mesh(path, **kw):
    from .meshio import read_msh
    return read_msh(path, **kw)
