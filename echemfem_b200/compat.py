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
    shim.__path__ = []                                  # a package: `from firedrake.petsc import PETSc`
    sys.modules["firedrake"] = shim
    pkg = types.ModuleType("echemfem")
    for n in ("EchemSolver", "TransientEchemSolver", "IntervalBoundaryLayerMesh", "RectangleBoundaryLayerMesh"):
        setattr(pkg, n, getattr(echemfem_b200, n))
    sys.modules["echemfem"] = pkg
    sys.modules["echemflow"] = pkg
    _stub_third_party()
    return True


class _Anything:
    """Permissive stand-in for plotting objects: every attribute is callable and returns another one."""

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))

    def __getitem__(self, i):
        return _Anything()


def _stub_third_party():
    """Single-process stand-ins for the MPI / PETSc / plotting imports at the top of the reference's examples
    (`from mpi4py import MPI`, `from petsc4py import PETSc`, `import matplotlib.pyplot as plt`) -- installed only
    when the real package is missing.  One process per GPU: rank 0 of 1 unless torch.distributed says otherwise."""
    def missing(name):
        try:
            importlib.import_module(name)
            return False
        except Exception:
            return True

    if missing("mpi4py"):
        class _Comm:
            rank, size = 0, 1

            def Get_rank(self):
                return int(__import__("os").environ.get("RANK", "0"))

            def Get_size(self):
                return int(__import__("os").environ.get("WORLD_SIZE", "1"))

            # under torchrun (one process per GPU) the collectives go through torch.distributed; a stub that
            # reports WORLD_SIZE ranks but reduces nothing would make reference scripts print wrong sums silently
            @staticmethod
            def _dist():
                if int(__import__("os").environ.get("WORLD_SIZE", "1")) == 1:
                    return None
                import torch.distributed as dist
                if not dist.is_initialized():
                    raise RuntimeError("mpi4py stand-in: WORLD_SIZE > 1 but torch.distributed is not initialised")
                return dist

            def Barrier(self):
                d = self._dist()
                if d is not None:
                    d.barrier()

            barrier = Barrier

            def bcast(self, obj, root=0):
                d = self._dist()
                if d is None:
                    return obj
                box = [obj]
                d.broadcast_object_list(box, src=root)
                return box[0]

            def allreduce(self, v, op=None):
                d = self._dist()
                if d is None:
                    return v
                vals = [None] * d.get_world_size()
                d.all_gather_object(vals, v)
                if op == "max":
                    return max(vals)
                if op == "min":
                    return min(vals)
                out = vals[0]
                for x in vals[1:]:
                    out = out + x
                return out
        mpi = types.ModuleType("mpi4py")
        MPI = types.ModuleType("mpi4py.MPI")
        MPI.Comm = _Comm
        MPI.COMM_WORLD = _Comm()
        MPI.COMM_SELF = _Comm()
        MPI.SUM, MPI.MAX, MPI.MIN = "sum", "max", "min"
        MPI.Wtime = __import__("time").time
        mpi.MPI = MPI
        sys.modules["mpi4py"] = mpi
        sys.modules["mpi4py.MPI"] = MPI
    if missing("petsc4py"):
        class _Sys:
            @staticmethod
            def Print(*a, **k):
                k.pop("comm", None)
                print(*a, **k)

            @staticmethod
            def popErrorHandler():
                pass

        class _Options(dict):
            def getInt(self, k, default=None):
                return int(self.get(k, default))

            def getReal(self, k, default=None):
                return float(self.get(k, default))

            def getString(self, k, default=None):
                return str(self.get(k, default))

            def getBool(self, k, default=False):
                return bool(self.get(k, default))

            def hasName(self, k):
                return k in self

            def setValue(self, k, v):
                self[k] = v

        class _Log:
            @staticmethod
            def begin():
                pass

            @staticmethod
            def Stage(name):
                return _Anything()

            @staticmethod
            def Event(name):
                return _Anything()
        pk = types.ModuleType("petsc4py")
        PETSc = types.ModuleType("petsc4py.PETSc")
        PETSc.Sys, PETSc.Options, PETSc.Log = _Sys, _Options, _Log
        PETSc.COMM_WORLD = None
        pk.PETSc = PETSc
        pk.init = lambda *a, **k: None
        sys.modules["petsc4py"] = pk
        sys.modules["petsc4py.PETSc"] = PETSc
    if missing("matplotlib"):
        mpl = types.ModuleType("matplotlib")
        mpl.__path__ = []
        mpl.use = lambda *a, **k: None
        mpl.rcParams = {}
        for sub in ("pyplot", "gridspec", "ticker", "cm", "colors"):
            m = types.ModuleType("matplotlib." + sub)
            m.__getattr__ = lambda name: _Anything()
            setattr(mpl, sub, m)
            sys.modules["matplotlib." + sub] = m
        sys.modules["matplotlib"] = mpl
    if "firedrake.petsc" not in sys.modules and "firedrake" in sys.modules:
        fp = types.ModuleType("firedrake.petsc")
        fp.PETSc = importlib.import_module("petsc4py.PETSc") if not missing("petsc4py.PETSc") else sys.modules["petsc4py.PETSc"]
        sys.modules["firedrake.petsc"] = fp
        sys.modules["firedrake"].petsc = fp


installed = install()
