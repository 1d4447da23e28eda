"""Run the reference's scripts unchanged: `import echemfem_b200.compat` before them.

The reference's examples and tests start with `from firedrake import *` and `from echemfem import
EchemSolver`.  When those packages are not importable, this module installs
`echemfem_b200.fd` as `firedrake` and `echemfem_b200` as `echemfem` (and the stale `echemflow` name
used by examples/bortels_threeion_nondim.py:2), so such a script drives the B200 path as is.
If a real Firedrake is importable nothing is replaced.
"""
import importlib
import sys
import types


def install(force=False):
    try:
        if not force:
            importlib.import_module("firedrake")
            return False
    except Exception:
        pass
    from . import fd
    import echemfem_b200
    shim = types.ModuleType("firedrake")
    names = [n for n in dir(fd) if not n.startswith("_")]
    for n in names:
        setattr(shim, n, getattr(fd, n))
    shim.__all__ = names
    sys.modules["firedrake"] = shim
    pkg = types.ModuleType("echemfem")
    for n in ("EchemSolver", "TransientEchemSolver", "IntervalBoundaryLayerMesh", "RectangleBoundaryLayerMesh"):
        setattr(pkg, n, getattr(echemfem_b200, n))
    sys.modules["echemfem"] = pkg
    sys.modules["echemflow"] = pkg
    return True


installed = install()
