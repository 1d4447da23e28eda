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
from .function import (Function, FunctionSpace, MixedFunctionSpace, VectorFunctionSpace, VectorFunction,  # noqa: F401
                       errornorm)


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


def IntervalMesh(ncells, length_or_left, right=None, **kw):
    """Uniform interval mesh; boundary ids 1 (left) and 2 (right), as Firedrake's."""
    import numpy as np
    left, right = (0.0, float(length_or_left)) if right is None else (float(length_or_left), float(right))
    return TensorMesh([np.linspace(left, right, int(ncells) + 1)], name="interval")


def UnitIntervalMesh(ncells, **kw):
    return IntervalMesh(ncells, 1.0)


def MeshHierarchy(mesh, refinement_levels, **kw):
    """Uniform refinements of a tensor-product mesh (every cell split in two per direction and level);
    `hierarchy[-1]` is what the reference's scripts take (examples/mms_scaling.py:72-76)."""
    import numpy as np
    if not hasattr(mesh, "axes"):
        raise NotImplementedError("MeshHierarchy needs a tensor-product mesh")
    out = [mesh]
    axes = [np.asarray(a, dtype=float) for a in mesh.axes]
    for _ in range(int(refinement_levels)):
        axes = [np.sort(np.concatenate([a, 0.5 * (a[1:] + a[:-1])])) for a in axes]
        out.append(TensorMesh(axes, name=mesh.name, layers=mesh.layers, marker_of_facet=mesh._marker_of_facet))
    return out


def ExtrudedMeshHierarchy(base_hierarchy, height, base_layer=1, refinement_ratio=2, **kw):
    """Extrusion of every level of a hierarchy: `base_layer` layers on the coarsest level, `refinement_ratio`
    times more per level, total height `height`."""
    out = []
    for lvl, m in enumerate(base_hierarchy):
        layers = int(base_layer) * int(refinement_ratio) ** lvl
        out.append(ExtrudedMesh(m, layers, float(height) / layers))
    return out


# logging knobs of `from firedrake import *` scripts (no-ops here)
DEBUG, INFO, WARNING = 10, 20, 30


def set_log_level(level):
    pass


class AssembledPC:
    """Base class of the reference's custom PETSc preconditioners (examples/mms_scaling.py:24-62).  Python-level
    PETSc preconditioners have no counterpart on the device path: the classes can be DEFINED (so that such scripts
    import), but instantiating one -- which is what selecting `pc_type: python` does -- raises instead of silently
    solving with a different preconditioner."""

    def __init__(self, *a, **k):
        raise NotImplementedError("Python-level PETSc preconditioners (%s) have no counterpart on the device path; "
                                  "use the option sets of init_solver_parameters ('lu', 'ilu', 'gmg')"
                                  % type(self).__name__)


class PMGPC(AssembledPC):
    pass


class PCBase(AssembledPC):
    pass
