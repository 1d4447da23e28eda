"""Benchmark workloads of BASELINE.json, written against the EchemSolver contract (product side).

`MMSSolver` is the manufactured electroneutral Nernst-Planck problem of the reference's scaling study
(examples/mms_scaling.py:66-155: fields (C1, Phi), C2 eliminated, D = (0.5e-5, 1e-5), z = (2, -2), parabolic
channel velocity, Dirichlet data from the exact fields on all sides) on a quadrilateral mesh -- the same class as
the reference's convergence test (tests/test_advection_diffusion_migration.py:11-63), whose coefficients it
takes as parameters.  bench.py, __graft_entry__ and the tests all build the benchmark problem from here.
"""
from .fd import (UnitSquareMesh, SpatialCoordinate, Constant, as_vector, cos, sin, div, grad)
from .solver import EchemSolver

CFG2_D = (0.5e-5, 1.0e-5)                      # examples/mms_scaling.py:90-93
CFG2_N = {1: 1408, 2: 944, 3: 704}             # cells per direction giving ~16 M dofs at Q1 / Q2 / Q3


class MMSSolver(EchemSolver):
    def __init__(self, N, vavg, family="DG", p=1, D1=0.5, D2=1.0, mesh=None):
        if mesh is None:
            mesh = UnitSquareMesh(N, N, quadrilateral=True)
        xs = SpatialCoordinate(mesh)
        x, y = xs[0], xs[1]
        C1ex = cos(x) + sin(y) + 3
        C2ex = C1ex
        Uex = sin(x) + cos(y) + 3
        self.C1ex, self.C2ex, self.Uex = C1ex, C2ex, Uex
        z1, z2 = 2.0, -2.0

        def f(C):
            f1 = div((self.vel - D1 * z1 * grad(Uex)) * C1ex) - div(D1 * grad(C1ex))
            f2 = div((self.vel - D2 * z2 * grad(Uex)) * C2ex) - div(D2 * grad(C2ex))
            return [f1, f2]

        conc_params = [{"name": "C1", "diffusion coefficient": D1, "z": z1, "bulk": C1ex},
                       {"name": "C2", "diffusion coefficient": D2, "z": z2, "bulk": C2ex}]
        physical_params = {"flow": ["diffusion", "advection", "migration", "electroneutrality"],
                           "F": 1.0, "R": 1.0, "T": 1.0, "U_app": Uex, "bulk reaction": f,
                           "v_avg": Constant(vavg)}
        super().__init__(conc_params, physical_params, mesh, family=family, p=p)

    def set_boundary_markers(self):
        self.boundary_markers = {"applied": (1, 2, 3, 4,), "bulk dirichlet": (1, 2, 3, 4,)}

    def set_velocity(self):
        xs = SpatialCoordinate(self.mesh)
        y = xs[1]
        h = 1.0
        comps = [6. * self.physical_params["v_avg"] / h**2 * y * (h - y)] + [Constant(0.)] * (len(xs) - 1)
        self.vel = as_vector(comps)


def cfg2_solver(N=None, p=1, mesh=None):
    """BASELINE.json configs[1]: the MMS problem with the scaling study's diffusion coefficients."""
    return MMSSolver(N if N is not None else CFG2_N[p], 1.0, D1=CFG2_D[0], D2=CFG2_D[1], p=p, mesh=mesh)


def cfg2_state(node_coords):
    """Benchmark state: nodal interpolant of the exact fields + 1e-3 sin(7x + 3y), both fields."""
    import numpy as np
    x = node_coords
    pert = 1e-3 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    return np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3 + pert, np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3 + pert])
