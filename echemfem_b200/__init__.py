"""echemfem_b200 -- B200-native Newton step behind the EchemFEM solver contract.

Public names mirror `echemfem/__init__.py:1-4` of the reference for the parts
that are on the accelerated path.
"""
from .solver import EchemSolver, TransientEchemSolver  # noqa: F401
from .meshes import IntervalBoundaryLayerMesh, RectangleBoundaryLayerMesh  # noqa: F401  (utility_meshes.py:6)
