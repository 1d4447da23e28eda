"""The reference's test problems written against the product's EchemSolver contract.

Same classes as the reference's tests (tests/test_advection_diffusion_migration.py:6-64
etc.), with `from firedrake import *` replaced by the package's own namespace.
"""
from echemfem_b200.fd import *  # noqa: F401,F403
from echemfem_b200 import EchemSolver


class AdvectionDiffusionMigrationSolver(EchemSolver):
    def __init__(self, N, vavg, family="DG", p=1, D1=0.5, D2=1.0, mesh=None):
        if mesh is None:
            mesh = UnitSquareMesh(N, N, quadrilateral=True)
        conc_params = []
        x, y = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        C1ex = cos(x) + sin(y) + 3
        C2ex = C1ex
        Uex = sin(x) + cos(y) + 3
        self.C1ex = C1ex
        self.C2ex = C2ex
        self.Uex = Uex

        def f(C):
            z1 = 2.0
            z2 = -2.0
            f1 = div((self.vel - D1 * z1 * grad(Uex)) * C1ex) - div(D1 * grad(C1ex))
            f2 = div((self.vel - D2 * z2 * grad(Uex)) * C2ex) - div(D2 * grad(C2ex))
            return [f1, f2]

        conc_params.append({"name": "C1", "diffusion coefficient": D1, "z": 2., "bulk": C1ex})
        conc_params.append({"name": "C2", "diffusion coefficient": D2, "z": -2., "bulk": C2ex})
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
