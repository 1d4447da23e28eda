"""The reference's test problems written against the product's EchemSolver contract.

Same classes as the reference's tests (tests/test_advection_diffusion_migration.py:6-64
etc.), with `from firedrake import *` replaced by the package's own namespace.
"""
from echemfem_b200.fd import *  # noqa: F401,F403
from echemfem_b200 import EchemSolver


from echemfem_b200.workloads import MMSSolver as AdvectionDiffusionMigrationSolver  # noqa: E402,F401
# (tests/test_advection_diffusion_migration.py:11-63 of the reference; the class lives in the package because it
# is also the benchmark workload of BASELINE.json configs[1])


def bortels_parameters():
    """Non-dimensional three-ion parameters of examples/bortels_threeion_nondim.py:28-47."""
    D_Cu, J0, C_Cu = 7.2e-10, 30., 10.
    D_SO4, C_SO4 = 10.65e-10, 1010.
    D_H, C_H = 93.12e-10, 2000.
    F, R, T = 96485.3329, 8.3144598, 273.15 + 25.
    U_app, v_avg, L = 0.006, 0.03, 0.01
    return {"D_Cu_hat": D_Cu / L / v_avg, "D_SO4_hat": D_SO4 / L / v_avg, "D_H_hat": D_H / L / v_avg,
            "J0_hat": J0 / v_avg / C_Cu / F, "U_app_hat": F / R / T * U_app,
            "C_SO4": C_SO4 / C_Cu, "C_H": C_H / C_Cu}


class BortelsSolver(EchemSolver):
    """examples/bortels_threeion_nondim.py:24-135 of the reference (BASELINE.json configs[0])."""

    def __init__(self, mesh, U_app_factor=1.0, current_shape=1.0):
        P = bortels_parameters()
        conc_params = [
            {"name": "Cu", "diffusion coefficient": P["D_Cu_hat"], "z": 2., "bulk": 1., "C_ND": 1.},
            {"name": "SO4", "diffusion coefficient": P["D_SO4_hat"], "z": -2., "bulk": 1., "C_ND": P["C_SO4"]},
            {"name": "H", "diffusion coefficient": P["D_H_hat"], "z": 1., "bulk": 1., "C_ND": P["C_H"]},
        ]
        physical_params = {"flow": ["advection", "diffusion", "migration", "electroneutrality"],
                           "F": 1.0, "R": 1.0, "T": 1.0, "U_app": P["U_app_hat"] * U_app_factor, "v_avg": 1.0}

        def reaction(u, V):
            C = u[0]
            U = u[2]
            C_b = conc_params[0]["bulk"]
            F, R, T = physical_params["F"], physical_params["R"], physical_params["T"]
            eta = V - U
            return current_shape * P["J0_hat"] * (exp(F / R / T * (eta)) - (C / C_b) * exp(-F / R / T * (eta)))

        def reaction_cathode(u):
            return reaction(u, Constant(0))

        def reaction_anode(u):
            return reaction(u, Constant(physical_params["U_app"]))

        echem_params = [
            {"reaction": reaction_cathode, "electrons": 2, "stoichiometry": {"Cu": 1}, "boundary": "cathode"},
            {"reaction": reaction_anode, "electrons": 2, "stoichiometry": {"Cu": 1}, "boundary": "anode"},
        ]
        super().__init__(conc_params, physical_params, mesh, echem_params=echem_params)

    def set_boundary_markers(self):
        self.boundary_markers = {"inlet": (10,), "outlet": (11,), "anode": (12,), "cathode": (13,)}

    def set_velocity(self):
        y = SpatialCoordinate(self.mesh)[1]
        h = 1.0
        self.vel = as_vector((6. * self.physical_params["v_avg"] / h**2 * y * (h - y), Constant(0.)))


class AdvectionDiffusionSolver(EchemSolver):
    """tests/test_SUPG.py:6-66 of the reference: CG advection-diffusion of two species with SUPG."""

    def __init__(self, N, D, SUPG=True, family="CG", mesh=None):
        if mesh is None:
            mesh = UnitSquareMesh(N, N, quadrilateral=True)
        x, y = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        C1ex = sin(x) + cos(y)
        C2ex = cos(x) + sin(y)
        self.C1ex = C1ex
        self.C2ex = C2ex

        def f(C):
            f1 = div(self.vel * C1ex - D * grad(C1ex))
            f2 = div(self.vel * C2ex - D * grad(C2ex))
            return [f1, f2]

        conc_params = [{"name": "C1", "diffusion coefficient": D, "bulk": C1ex},
                       {"name": "C2", "diffusion coefficient": D, "bulk": C2ex}]
        physical_params = {"flow": ["diffusion", "advection"], "bulk reaction": f, "v_avg": 1.0}
        super().__init__(conc_params, physical_params, mesh, family=family, SUPG=SUPG)

    def set_boundary_markers(self):
        self.boundary_markers = {"inlet": (1, 2, 3, 4), "outlet": (1, 2, 3, 4), "bulk dirichlet": (1, 2, 3, 4)}

    def set_velocity(self):
        xs = SpatialCoordinate(self.mesh)
        y = xs[1]
        h = 1.0
        x_vel = 6. * self.physical_params["v_avg"] / h**2 * y * (h - y)
        self.vel = as_vector([x_vel] + [Constant(0.)] * (len(xs) - 1))


class DiffusionMigrationPoissonSolver(EchemSolver):
    """tests/test_diffusion_migration_poisson.py:6-57 of the reference: Nernst-Planck + true Poisson."""

    def __init__(self, N, family="DG", variant=None):
        mesh = UnitSquareMesh(N, N, quadrilateral=True)
        x, y = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        C1ex = cos(x) + sin(y) + 3
        C2ex = cos(x) - sin(y) + 3
        Uex = sin(x) + cos(y) + 3
        self.C1ex, self.C2ex, self.Uex = C1ex, C2ex, Uex
        self.variant = variant        # Poisson branches of solver.py:2085-2093, 2219-2232 (see mms_cases.poisson_problem)

        def f(C):
            D1, D2, z1, z2, K = 0.5, 1.0, 2.0, -2.0, 2e-1
            f1 = div((- D1 * z1 * grad(Uex)) * C1ex) - div(D1 * grad(C1ex))
            f2 = div((- D2 * z2 * grad(Uex)) * C2ex) - div(D2 * grad(C2ex))
            f3 = div((-K * grad(Uex))) - z1 * C1ex - z2 * C2ex
            return [f1, f2, f3]

        conc_params = [{"name": "C1", "diffusion coefficient": 0.5, "z": 2., "bulk": C1ex},
                       {"name": "C2", "diffusion coefficient": 1.0, "z": -2., "bulk": C2ex}]
        physical_params = {"flow": ["diffusion", "migration", "poisson"], "F": 1.0, "R": 1.0, "T": 1.0,
                           "vacuum permittivity": 2.0, "relative permittivity": 1e-1, "U_app": Uex,
                           "bulk reaction": f}
        if variant == "robin":
            physical_params["gap capacitance"] = 0.7
        elif variant == "applied liquid current":
            physical_params["applied current density"] = 0.3
        elif variant == "liquid conductivity":
            del physical_params["vacuum permittivity"], physical_params["relative permittivity"]
            physical_params["liquid conductivity"] = 0.4
        super().__init__(conc_params, physical_params, mesh, family=family)

    def set_boundary_markers(self):
        self.boundary_markers = {"bulk dirichlet": (1, 2, 3, 4,), "applied": (1, 2, 3, 4,)}
        if self.variant == "robin":
            self.boundary_markers.update({"applied": (1, 3, 4), "robin": (2,)})
        elif self.variant == "applied liquid current":
            self.boundary_markers.update({"applied": (1, 3, 4), "applied liquid current": (2,)})


from echemfem_b200 import TransientEchemSolver  # noqa: E402


class HeatEquationSolver(TransientEchemSolver):
    """tests/test_heat_equation.py:7-42 of the reference: transient diffusion of two species, backward Euler."""

    def __init__(self, N, family="DG"):
        mesh = UnitSquareMesh(N, N, quadrilateral=True)
        self.time = Constant(0)
        t = self.time
        x, y = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        C1ex = (sin(x) + cos(y)) * exp(t)
        C2ex = (cos(x) + sin(y)) * exp(-t)
        self.C1ex, self.C2ex = C1ex, C2ex

        def f(C):
            f1 = C1ex - div(grad(C1ex))
            f2 = -C2ex - div(grad(C2ex))
            return [f1, f2]

        conc_params = [{"name": "C1", "diffusion coefficient": 1.0, "bulk": C1ex},
                       {"name": "C2", "diffusion coefficient": 1.0, "bulk": C2ex}]
        physical_params = {"flow": ["diffusion"], "bulk reaction": f}
        super().__init__(conc_params, physical_params, mesh, family=family)

    def set_boundary_markers(self):
        self.boundary_markers = {"bulk dirichlet": (1, 2, 3, 4,)}


class PorousSolver(EchemSolver):
    """tests/test_advection_diffusion_migration_porous.py:6-70 of the reference: porous electrode model,
    liquid + solid potentials, Bruggeman effective coefficients."""

    def __init__(self, N, vavg, solid_current=False):
        mesh = UnitSquareMesh(N, N, quadrilateral=True)
        x, y = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        C1ex = cos(x) + sin(y) + 3
        C2ex = C1ex
        U1ex = sin(x) + cos(y) + 3
        U2ex = sin(x) + cos(y) + 3
        self.C1ex, self.C2ex, self.U1ex, self.U2ex = C1ex, C2ex, U1ex, U2ex
        self.solid_current = solid_current       # "applied solid current" on side 2 (solver.py:2219-2221)

        def f(C):
            D1, D2, z1, z2, K = 0.5, 1.0, 2.0, -2.0, 1.0
            f1 = div((self.vel - self.effective_diffusion(D1) * z1 * grad(U1ex)) * C1ex) - \
                div(self.effective_diffusion(D1) * grad(C1ex))
            f2 = div((self.vel - self.effective_diffusion(D2) * z2 * grad(U1ex)) * C2ex) - \
                div(self.effective_diffusion(D2) * grad(C2ex))
            f3 = div(- self.effective_diffusion(K, phase="solid") * grad(U2ex))
            return [f1, f2, f3]

        conc_params = [{"name": "C1", "diffusion coefficient": 0.5, "z": 2., "bulk": C1ex},
                       {"name": "C2", "diffusion coefficient": 1.0, "z": -2., "bulk": C2ex}]
        physical_params = {"flow": ["diffusion", "advection", "migration", "electroneutrality", "porous"],
                           "F": 1.0, "R": 1.0, "T": 1.0, "U_app": U1ex, "bulk reaction": f, "v_avg": vavg,
                           "porosity": 0.5, "solid conductivity": 1., "specific surface area": 1.,
                           "standard potential": 0.}
        if solid_current:
            physical_params["applied current density"] = 0.25
        super().__init__(conc_params, physical_params, mesh)

    def set_boundary_markers(self):
        self.boundary_markers = {"bulk dirichlet": (1, 2, 3, 4,), "applied": (1, 2, 3, 4,),
                                 "liquid applied": (1, 2, 3, 4)}
        if self.solid_current:
            self.boundary_markers.update({"applied": (1, 3, 4), "applied solid current": (2,)})

    def set_velocity(self):
        y = SpatialCoordinate(self.mesh)[1]
        h = 1.0
        self.vel = as_vector((6. * self.physical_params["v_avg"] / h**2 * y * (h - y), Constant(0.)))


FS_SPECIES = [("A", 1.0, 1.0, 0.5), ("B", 0.5, -1.0, 0.4), ("N", 0.8, 0.0, 0.3)]      # name, D, z, solvated diameter
FS_HOMOG = [{"stoichiometry": {"A": -1, "B": -1, "N": 1}, "forward rate constant": 0.7,
             "backward rate constant": 0.3, "reference concentration": 2.0},
            {"stoichiometry": {"N": -2, "A": 1}, "forward rate constant": 0.05, "equilibrium constant": 4.0,
             "reaction order": {"N": 1}}]


class FiniteSizeSolver(EchemSolver):
    """Poisson-Nernst-Planck with a finite-size correction of the chemical potential (examples/BMCSL.py:33-84 in
    2D, dimensionless numbers) and homogeneous reactions by the law of mass action (homog_params,
    examples/carbonate_homog_params.py).  CG only (solver.py:139-142)."""

    def __init__(self, N, kind="BMCSL", p=1, homog=True):
        mesh = UnitSquareMesh(N, N, quadrilateral=True)
        x, y = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        bulk = {"A": cos(x) + sin(y) + 3, "B": cos(x) - sin(y) + 3, "N": sin(x) * cos(y) + 2}
        conc_params = [{"name": n, "diffusion coefficient": D, "z": z, "solvated diameter": a, "bulk": bulk[n]}
                       for n, D, z, a in FS_SPECIES]
        flow = ["migration", "poisson"]
        flow += ["diffusion", "finite size"] if kind == "finite size" else ["diffusion finite size_" + kind]
        physical_params = {"flow": flow, "F": 1.0, "R": 1.0, "T": 1.0, "vacuum permittivity": 2.0,
                           "relative permittivity": 1e-1, "U_app": sin(x) + cos(y) + 3, "Avogadro constant": 0.1}
        super().__init__(conc_params, physical_params, mesh, family="CG", p=p,
                         homog_params=[dict(r) for r in FS_HOMOG] if homog else [])

    def set_boundary_markers(self):
        self.boundary_markers = {"bulk dirichlet": (1, 2, 3, 4,), "applied": (1, 2, 3, 4,)}


GUPTA_SPECIES = [("CO2", 19.1e-10, 34.2, 0), ("HCO3", 9.23e-10, 499., -1), ("CO3", 11.9e-10, 7.6e-1, -2),
                 ("OH", 52.7e-10, 3.3e-4, -1), ("K", 19.6e-10, 499. + 7.6e-1 + 3.3e-4, 1)]
GUPTA_HOMOG = [{"stoichiometry": {"CO2": -1, "OH": -1, "HCO3": 1}, "forward rate constant": 5.93,
                "backward rate constant": 1.34e-4},
               {"stoichiometry": {"HCO3": -1, "OH": -1, "CO3": 1}, "forward rate constant": 1e5,
                "backward rate constant": 2.15e4}]
GUPTA_CURRENTS = [(0.1, {"CO2": -1, "OH": 1}, 2), (0.05, {"CO2": -1, "OH": 2}, 2), (0.25, {"CO2": -1, "OH": 8}, 8),
                  (0.2, {"CO2": -2, "OH": 12}, 12), (0.4, {"OH": 2}, 2)]
GUPTA_LX, GUPTA_LY = 1e-2, 1e-3


class GuptaSolver(EchemSolver):
    """examples/gupta_advection_migration.py:17-140 of the reference (BASELINE.json configs[2]) on a small mesh:
    CG electroneutral Nernst-Planck, 5 species (K eliminated), two homogeneous reactions, constant-current
    electrode fluxes on the active part of the wall, shear flow, inlet / outlet / bulk boundaries."""

    def __init__(self, nx=6, ny=4, n_layer=3, SUPG=False):
        from echemfem_b200 import RectangleBoundaryLayerMesh
        Lx, Ly = GUPTA_LX, GUPTA_LY
        mesh = RectangleBoundaryLayerMesh(nx, ny, Lx, Ly, n_layer, 1e-6, boundary=(3,))
        x, y = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        active = conditional(And(x >= 1e-3, x < Lx - 1e-3), 1., 0.)
        conc_params = [{"name": n, "diffusion coefficient": D, "bulk": b, "z": z} for n, D, b, z in GUPTA_SPECIES]
        physical_params = {"flow": ["diffusion", "electroneutrality", "migration", "advection"],
                           "F": 96485.3329, "R": 8.3144598, "T": 273.15 + 25.}

        def current(cef):
            def curr(u):
                return cef * 50. * active
            return curr
        echem_params = [{"reaction": current(c), "stoichiometry": dict(st), "electrons": ne, "boundary": "electrode"}
                        for c, st, ne in GUPTA_CURRENTS]
        super().__init__(conc_params, physical_params, mesh, family="CG", echem_params=echem_params,
                         homog_params=[dict(r) for r in GUPTA_HOMOG], SUPG=SUPG)

    def set_boundary_markers(self):
        self.boundary_markers = {"inlet": (1,), "outlet": (2,), "bulk": (4,), "electrode": (3,)}

    def set_velocity(self):
        y = SpatialCoordinate(self.mesh)[1]
        self.vel = as_vector([1.91 * y, Constant(0)])


class Bortels3DSolver(BortelsSolver):
    """examples/bortels_threeion_extruded_3D_nondim.py:74-222 of the reference (BASELINE.json configs[3]): the plane
    channel extruded in z (DQ1 x DG1 hexes), exchange current density parabolic in z, side walls keep the ids of
    the plane mesh, top and bottom are never integrated (solver.py:221-228)."""

    def __init__(self, mesh, hz):
        z = SpatialCoordinate(mesh)[2]
        shape = 3.0 / 5.0 * (2 - (z - 0.5 * hz) ** 2 / (0.5 * hz) ** 2)
        super().__init__(mesh, current_shape=shape)

    def set_velocity(self):
        y = SpatialCoordinate(self.mesh)[1]
        self.vel = as_vector((6. * self.physical_params["v_avg"] * y * (1.0 - y), Constant(0.), Constant(0.)))


CL_SPECIES = [("CO2", 1.9, 0.0, 0.9, 1.3, 1.1), ("OH", 5.3, -1.0, 0.4, 0.8, None), ("H", 9.3, 1.0, 0.2, 1.7, None),
              ("CO3", 0.9, -2.0, 0.3, 0.6, None), ("HCO3", 1.2, -1.0, 0.8, 0.9, None),
              ("K", 2.0, 1.0, 1.6, 1.1, None)]          # name, D, z, bulk, mass transfer coefficient, gas value
CL_PARAMS = {"F": 1.0, "R": 1.0, "T": 1.0, "porosity": 0.6, "saturation": 0.64, "specific surface area": 1.5,
             "solid conductivity": 2.2}
CL_RATES = {"k1f": 0.7, "k1b": 0.2, "k2f": 1.1, "k2b": 0.4, "kwf": 0.05, "kwb": 0.9}
CL_ECHEM = [(0.02, 0.5, 0.11, {"CO2": 1, "OH": -6}, 6), (0.03, 0.4, -0.05, {"OH": -2}, 2)]   # i0, alpha, U0, nu, n


class CatalystLayerSolver(EchemSolver):
    """examples/catalyst_layer_Cu_full.py:20-256 of the reference (BASELINE.json configs[4]) with O(1) numbers on
    a 2D strip: porous electrode, six species with the explicit electroneutrality constraint ("electroneutrality
    full": shifted test functions, solver.py:411-414, 455-461), liquid + solid potential, homogeneous carbonate
    reactions, volumetric Tafel kinetics depending on c_H and both potentials, gas-side Dirichlet value of CO2,
    Robin exchange with the bulk electrolyte."""

    def __init__(self, N, p=1):
        mesh = RectangleMesh(N, max(N // 2, 1), 1.0, 0.5, quadrilateral=True)
        P, K = CL_PARAMS, CL_RATES
        epsS = P["porosity"] * P["saturation"]

        def bulk_reaction(y):
            CO2, OH, H, CO3, HCO3 = y[0], y[1], y[2], y[3], y[4]
            r1 = K["k1f"] * CO2 * OH - K["k1b"] * HCO3
            r2 = K["k2f"] * HCO3 * OH - K["k2b"] * CO3
            rw = K["kwf"] - K["kwb"] * H * OH
            return [epsS * (-r1), epsS * (-r1 - r2 + rw), epsS * rw, epsS * r2, epsS * (r1 - r2), 0.]

        def tafel(i0, alpha, U0):
            def r(u):
                eta = u[7] - u[6] - (U0 + P["R"] * P["T"] / P["F"] * ln(u[2] / 0.2))
                return -i0 * exp(-alpha * P["F"] / (P["R"] * P["T"]) * eta)
            return r
        conc_params = []
        for n, D, z, b, kmt, gas in CL_SPECIES:
            cp = {"name": n, "diffusion coefficient": D, "z": z, "bulk": b, "mass transfer coefficient": kmt}
            if gas is not None:
                cp["gas"] = gas
            conc_params.append(cp)
        physical_params = dict(P, flow=["diffusion", "migration", "electroneutrality full", "porous"], U_app=-0.3)
        physical_params["bulk reaction"] = bulk_reaction
        echem_params = [{"reaction": tafel(i0, al, U0), "electrons": ne, "stoichiometry": dict(nu)}
                        for i0, al, U0, nu, ne in CL_ECHEM]
        super().__init__(conc_params, physical_params, mesh, echem_params=echem_params, family="CG", p=p)

    def set_boundary_markers(self):
        self.boundary_markers = {"applied": (2,), "gas": (2,), "bulk": (1,)}


class BMCSL1DSolver(EchemSolver):
    """examples/BMCSL.py:33-84 of the reference with O(1) numbers: 1D Poisson-Nernst-Planck with the BMCSL
    finite-size correction on an interval with a boundary layer at the charged wall (CG P2): Dirichlet bulk /
    zero potential on the left, surface charge density (Poisson Neumann) on the right."""

    def __init__(self, p=2, sigma=-0.3):
        from echemfem_b200 import IntervalBoundaryLayerMesh
        mesh = IntervalBoundaryLayerMesh(6, 1.0, 4, 0.1, boundary=(2,))
        conc_params = [{"name": "Cl", "diffusion coefficient": 2.0, "bulk": 1.0, "z": -1, "solvated diameter": 0.0},
                       {"name": "Li", "diffusion coefficient": 1.0, "bulk": 0.9, "z": 1, "solvated diameter": 0.76},
                       {"name": "Cs", "diffusion coefficient": 2.0, "bulk": 0.1, "z": 1, "solvated diameter": 0.66}]
        physical_params = {"flow": ["migration", "poisson", "diffusion finite size_BMCSL"], "F": 1.0, "R": 1.0,
                           "T": 1.0, "U_app": Constant(0.0), "vacuum permittivity": 0.05, "relative permittivity": 1.0,
                           "Avogadro constant": 0.2, "surface charge density": Constant(sigma)}
        super().__init__(conc_params, physical_params, mesh, family="CG", p=p)

    def set_boundary_markers(self):
        self.boundary_markers = {"bulk dirichlet": (1,), "applied": (1,), "poisson neumann": (2,)}


class DiffusionFiniteSizeSolver(EchemSolver):
    """tests/test_diffusion_finite_size.py:6-46 of the reference: diffusion with the steric ("finite size") flux."""

    def __init__(self, N, family="DG"):
        mesh = UnitSquareMesh(N, N, quadrilateral=True)
        x, y = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        C1ex = sin(x) + cos(y)
        C2ex = cos(x) + sin(y)
        self.C1ex, self.C2ex = C1ex, C2ex

        def f(C):
            num = grad(C1ex) + grad(C2ex)
            den = 1.0 - C1ex - C2ex
            f1 = - div(grad(C1ex) + C1ex * num / den)
            f2 = - div(grad(C2ex) + C2ex * num / den)
            return [f1, f2]
        conc_params = [{"name": "C1", "diffusion coefficient": 1.0, "bulk": C1ex, "solvated diameter": 1.0},
                       {"name": "C2", "diffusion coefficient": 1.0, "bulk": C2ex, "solvated diameter": 1.0}]
        physical_params = {"flow": ["diffusion", "finite size"], "vacuum permittivity": 1.0,
                           "relative permittivity": 1.0, "Avogadro constant": 1.0, "bulk reaction": f}
        super().__init__(conc_params, physical_params, mesh, family=family)

    def set_boundary_markers(self):
        self.boundary_markers = {"bulk dirichlet": (1, 2, 3, 4)}


class DiffusionCylindricalSolver(EchemSolver):
    """tests/test_diffusion_cylindrical.py:6-41 of the reference: axisymmetric diffusion (r weight in every measure)."""

    def __init__(self, N, family="DG"):
        mesh = UnitSquareMesh(N, N, quadrilateral=True)
        r, z = SpatialCoordinate(mesh)[0], SpatialCoordinate(mesh)[1]
        C1ex = sin(z) + cos(r)
        C2ex = cos(z) * cos(r) + sin(z) * cos(r)
        self.C1ex, self.C2ex = C1ex, C2ex

        def f(C):
            f1 = sin(r) / r + cos(r) + sin(z)
            f2 = (r * cos(r) + sin(r)) * (cos(z) + sin(z)) / r + cos(r) * (cos(z) + sin(z))
            return [f1, f2]
        conc_params = [{"name": "C1", "diffusion coefficient": 1.0, "bulk": C1ex},
                       {"name": "C2", "diffusion coefficient": 1.0, "bulk": C2ex}]
        physical_params = {"flow": ["diffusion"], "bulk reaction": f}
        super().__init__(conc_params, physical_params, mesh, family=family, cylindrical=True)

    def set_boundary_markers(self):
        self.boundary_markers = {"bulk dirichlet": (2, 3, 4)}


class ScalarAdvectionDiffusionSolver(EchemSolver):
    """tests/test_advection_diffusion.py:6-65 of the reference: two passive species, DG (upwind) or CG (no SUPG),
    on quadrilaterals or on hexes extruded from them."""

    def __init__(self, N, D, extruded=False, family="DG"):
        if extruded:
            mesh = ExtrudedMesh(UnitSquareMesh(N, N, quadrilateral=True), N, layer_height=1.0 / N)
        else:
            mesh = UnitSquareMesh(N, N, quadrilateral=True)
        xs = SpatialCoordinate(mesh)
        x, y = xs[0], xs[1]
        C1ex = sin(x) + cos(y)
        C2ex = cos(x) + sin(y)
        self.C1ex, self.C2ex = C1ex, C2ex

        def f(C):
            f1 = div(self.vel * C1ex - D * grad(C1ex))
            f2 = div(self.vel * C2ex - D * grad(C2ex))
            return [f1, f2]
        conc_params = [{"name": "C1", "diffusion coefficient": D, "bulk": C1ex},
                       {"name": "C2", "diffusion coefficient": D, "bulk": C2ex}]
        physical_params = {"flow": ["diffusion", "advection"], "bulk reaction": f, "v_avg": 1.0}
        super().__init__(conc_params, physical_params, mesh, family=family)

    def set_boundary_markers(self):
        self.boundary_markers = {"inlet": (1, 2, 3, 4), "outlet": (1, 2, 3, 4), "bulk dirichlet": (1, 2, 3, 4)}

    def set_velocity(self):
        y = SpatialCoordinate(self.mesh)[1]
        x_vel = 6. * self.physical_params["v_avg"] * y * (1.0 - y)
        if self.mesh.layers is not None:
            self.vel = as_vector((x_vel, Constant(0.), Constant(0.)))
        else:
            self.vel = as_vector((x_vel, Constant(0.)))
