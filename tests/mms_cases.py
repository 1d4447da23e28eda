"""Manufactured-solution cases of the reference's tests, as numpy closures for the oracle.

Restates tests/test_advection_diffusion_migration.py:6-64 of the reference
(fields C1 = C2 = cos x + sin y + 3, U = sin x + cos y + 3, Poiseuille
velocity, D = (0.5, 1), z = (2, -2), F = R = T = 1); the forcing terms
f_k = div((vel - D_k z_k grad U) C_k) - div(D_k grad C_k) are expanded by hand.
"""
import numpy as np


def C1ex(x):
    return np.cos(x[..., 0]) + np.sin(x[..., 1]) + 3.0


def Uex(x):
    return np.sin(x[..., 0]) + np.cos(x[..., 1]) + 3.0


def make_vel(vavg, d=2):
    def vel(x):
        v = np.zeros(x.shape, dtype=float)
        y = x[..., 1]
        v[..., 0] = 6.0 * vavg * y * (1.0 - y)
        return v
    return vel


def forcing(vavg, D1=0.5, D2=1.0, z1=2.0, z2=-2.0):
    def f(u, x):
        X, Y = x[..., 0], x[..., 1]
        C = np.cos(X) + np.sin(Y) + 3.0
        Cx, Cy = -np.sin(X), np.cos(Y)
        lapC = -np.cos(X) - np.sin(Y)
        Ux, Uy = np.cos(X), -np.sin(Y)
        lapU = -np.sin(X) - np.cos(Y)
        vx = 6.0 * vavg * Y * (1.0 - Y)
        out = []
        for D, z in ((D1, z1), (D2, z2)):
            out.append(vx * Cx - D * z * (Cx * Ux + Cy * Uy + C * lapU) - D * lapC)
        return out
    return f


def adm_problem(vavg, family="DG", p=1, D1=0.5, D2=1.0, variant="spectral", markers=None):
    """AdvectionDiffusionMigrationSolver of the reference test, for the oracle."""
    from oracle.forms import Problem
    conc_params = [
        {"name": "C1", "diffusion coefficient": D1, "z": 2.0, "bulk": C1ex},
        {"name": "C2", "diffusion coefficient": D2, "z": -2.0, "bulk": C1ex},
    ]
    physical_params = {
        "flow": ["diffusion", "advection", "migration", "electroneutrality"],
        "F": 1.0, "R": 1.0, "T": 1.0, "U_app": Uex,
        "bulk reaction": forcing(vavg, D1, D2),
    }
    if markers is None:
        markers = {"applied": (1, 2, 3, 4), "bulk dirichlet": (1, 2, 3, 4)}
    return Problem(conc_params, physical_params, markers, p=p, family=family,
                   vel=make_vel(vavg), variant=variant)


def bortels_problem(U_app_factor=1.0, hz=None):
    """examples/bortels_threeion_nondim.py for the oracle: Butler-Volmer electrodes as numpy closures."""
    from oracle.forms import Problem
    import product_cases
    P = product_cases.bortels_parameters()
    U_app = P["U_app_hat"] * U_app_factor
    conc_params = [
        {"name": "Cu", "diffusion coefficient": P["D_Cu_hat"], "z": 2., "bulk": 1., "C_ND": 1.},
        {"name": "SO4", "diffusion coefficient": P["D_SO4_hat"], "z": -2., "bulk": 1., "C_ND": P["C_SO4"]},
        {"name": "H", "diffusion coefficient": P["D_H_hat"], "z": 1., "bulk": 1., "C_ND": P["C_H"]},
    ]
    physical_params = {"flow": ["advection", "diffusion", "migration", "electroneutrality"],
                       "F": 1.0, "R": 1.0, "T": 1.0, "U_app": U_app}

    def reaction(V):
        def r(u, x):
            eta = V - u[2]
            shape = 1.0
            if hz is not None:            # examples/bortels_threeion_extruded_3D_nondim.py: parabolic in z
                shape = 3.0 / 5.0 * (2.0 - (x[..., 2] - 0.5 * hz) ** 2 / (0.5 * hz) ** 2)
            return shape * P["J0_hat"] * (np.exp(eta) - u[0] * np.exp(-eta))
        return r
    echem_params = [
        {"reaction": reaction(0.0), "electrons": 2, "stoichiometry": {"Cu": 1}, "boundary": "cathode"},
        {"reaction": reaction(U_app), "electrons": 2, "stoichiometry": {"Cu": 1}, "boundary": "anode"},
    ]
    markers = {"inlet": (10,), "outlet": (11,), "anode": (12,), "cathode": (13,)}
    return Problem(conc_params, physical_params, markers, echem_params=echem_params, vel=make_vel(1.0))


def bortels_oracle_mesh(n_in, n_el, n_out, n_y, layers=None, hz=None):
    from oracle.mesh import tensor_mesh
    from echemfem_b200.meshes import bortels_axes, bortels_markers
    x, y = bortels_axes(n_in, n_el, n_out, n_y)
    fn = bortels_markers()

    def marker(xv):
        return int(fn(xv[None], None)[0])
    if layers is None:
        return tensor_mesh([x, y], boundary_id_fn=marker)
    z = np.linspace(0.0, hz, layers + 1)

    def marker3(xv):
        """side facets: id of the plane edge below; bottom / top: 5 / 6 (never integrated)."""
        zc = xv[:, 2]
        if np.ptp(zc) < 1e-12:
            return 5 if abs(zc[0]) < 1e-12 else 6
        pts = np.unique(np.round(xv[:, :2], 14), axis=0)
        return int(fn(pts[None], None)[0])
    return tensor_mesh([x, y, z], boundary_id_fn=marker3)


def supg_C1ex(x):
    return np.sin(x[..., 0]) + np.cos(x[..., 1])


def supg_C2ex(x):
    return np.cos(x[..., 0]) + np.sin(x[..., 1])


def supg_problem(D, SUPG=True, family="CG"):
    """tests/test_SUPG.py:6-66 of the reference for the oracle (forcing expanded by hand, div vel = 0)."""
    from oracle.forms import Problem

    def f(u, x):
        X, Y = x[..., 0], x[..., 1]
        vx = 6.0 * Y * (1.0 - Y)
        return [vx * np.cos(X) + D * (np.sin(X) + np.cos(Y)), -vx * np.sin(X) + D * (np.cos(X) + np.sin(Y))]
    conc_params = [{"name": "C1", "diffusion coefficient": D, "z": 0.0, "bulk": supg_C1ex},
                   {"name": "C2", "diffusion coefficient": D, "z": 0.0, "bulk": supg_C2ex}]
    physical_params = {"flow": ["diffusion", "advection"], "bulk reaction": f}
    markers = {"inlet": (1, 2, 3, 4), "outlet": (1, 2, 3, 4), "bulk dirichlet": (1, 2, 3, 4)}
    return Problem(conc_params, physical_params, markers, family=family, vel=make_vel(1.0), SUPG=SUPG)


def poisson_C2ex(x):
    return np.cos(x[..., 0]) - np.sin(x[..., 1]) + 3.0


def poisson_problem(family="DG", variant=None):
    """tests/test_diffusion_migration_poisson.py:6-57 of the reference for the oracle.

    `variant` switches on one of the Poisson branches the reference's own tests never reach
    (echemfem/solver.py:2085-2093, 2219-2232): "robin" (gap capacitance on side 2), "applied liquid current"
    (applied current density on side 2), "liquid conductivity" (K_U given instead of the permittivities)."""
    from oracle.forms import Problem

    def f(u, x):
        X, Y = x[..., 0], x[..., 1]
        D1, D2, z1, z2, K = 0.5, 1.0, 2.0, -2.0, 2e-1
        C1 = np.cos(X) + np.sin(Y) + 3.0
        C2 = np.cos(X) - np.sin(Y) + 3.0
        g1 = (-np.sin(X), np.cos(Y))
        g2 = (-np.sin(X), -np.cos(Y))
        l1 = -np.cos(X) - np.sin(Y)
        l2 = -np.cos(X) + np.sin(Y)
        gU = (np.cos(X), -np.sin(Y))
        lU = -np.sin(X) - np.cos(Y)
        f1 = -D1 * z1 * (g1[0] * gU[0] + g1[1] * gU[1] + C1 * lU) - D1 * l1
        f2 = -D2 * z2 * (g2[0] * gU[0] + g2[1] * gU[1] + C2 * lU) - D2 * l2
        f3 = -K * lU - z1 * C1 - z2 * C2
        return [f1, f2, f3]
    conc_params = [{"name": "C1", "diffusion coefficient": 0.5, "z": 2.0, "bulk": C1ex},
                   {"name": "C2", "diffusion coefficient": 1.0, "z": -2.0, "bulk": poisson_C2ex}]
    physical_params = {"flow": ["diffusion", "migration", "poisson"], "F": 1.0, "R": 1.0, "T": 1.0,
                       "vacuum permittivity": 2.0, "relative permittivity": 1e-1, "U_app": Uex, "bulk reaction": f}
    markers = {"bulk dirichlet": (1, 2, 3, 4), "applied": (1, 2, 3, 4)}
    if variant == "robin":
        physical_params["gap capacitance"] = 0.7
        markers.update({"applied": (1, 3, 4), "robin": (2,)})
    elif variant == "applied liquid current":
        physical_params["applied current density"] = 0.3
        markers.update({"applied": (1, 3, 4), "applied liquid current": (2,)})
    elif variant == "liquid conductivity":
        del physical_params["vacuum permittivity"], physical_params["relative permittivity"]
        physical_params["liquid conductivity"] = 0.4
    elif variant is not None:
        raise ValueError(variant)
    return Problem(conc_params, physical_params, markers, family=family)


def heat_problem(t, dt, u_old, family="DG"):
    """tests/test_heat_equation.py:7-42 at time t with a backward-Euler step of size dt from u_old."""
    from oracle.forms import Problem

    def c1(x):
        return (np.sin(x[..., 0]) + np.cos(x[..., 1])) * np.exp(t)

    def c2(x):
        return (np.cos(x[..., 0]) + np.sin(x[..., 1])) * np.exp(-t)

    def f(u, x):
        return [2.0 * c1(x), 0.0 * c2(x)]          # C1ex - lap C1ex = 2 C1ex ; -C2ex - lap C2ex = 0
    conc_params = [{"name": "C1", "diffusion coefficient": 1.0, "z": 0.0, "bulk": c1},
                   {"name": "C2", "diffusion coefficient": 1.0, "z": 0.0, "bulk": c2}]
    prob = Problem(conc_params, {"flow": ["diffusion"], "bulk reaction": f}, {"bulk dirichlet": (1, 2, 3, 4)},
                   family=family)
    prob.transient = {"dt": dt, "u_old": u_old}
    return prob, c1, c2


def porous_problem(vavg, solid_current=False):
    """tests/test_advection_diffusion_migration_porous.py:6-70 of the reference for the oracle (DG).
    `solid_current`: side 2 carries "applied solid current" (solver.py:2219-2221) instead of the applied potential."""
    from oracle.forms import Problem
    eps = 0.5
    De = lambda D: D * eps ** 1.5
    Ks = 1.0 * (1.0 - eps) ** 1.5

    def f(u, x):
        X, Y = x[..., 0], x[..., 1]
        C = np.cos(X) + np.sin(Y) + 3.0
        Cx, Cy = -np.sin(X), np.cos(Y)
        lapC = -np.cos(X) - np.sin(Y)
        Ux, Uy = np.cos(X), -np.sin(Y)
        lapU = -np.sin(X) - np.cos(Y)
        vx = 6.0 * vavg * Y * (1.0 - Y)
        out = []
        for D, z in ((0.5, 2.0), (1.0, -2.0)):
            out.append(vx * Cx - De(D) * z * (Cx * Ux + Cy * Uy + C * lapU) - De(D) * lapC)
        out.append(-Ks * lapU)
        return out
    conc_params = [{"name": "C1", "diffusion coefficient": 0.5, "z": 2.0, "bulk": C1ex},
                   {"name": "C2", "diffusion coefficient": 1.0, "z": -2.0, "bulk": C1ex}]
    physical_params = {"flow": ["diffusion", "advection", "migration", "electroneutrality", "porous"],
                       "F": 1.0, "R": 1.0, "T": 1.0, "U_app": Uex, "bulk reaction": f, "porosity": eps,
                       "solid conductivity": 1.0, "specific surface area": 1.0}
    markers = {"bulk dirichlet": (1, 2, 3, 4), "applied": (1, 2, 3, 4), "liquid applied": (1, 2, 3, 4)}
    if solid_current:
        physical_params["applied current density"] = 0.25
        markers.update({"applied": (1, 3, 4), "applied solid current": (2,)})
    return Problem(conc_params, physical_params, markers, vel=make_vel(vavg))


def finite_size_problem(kind="BMCSL", p=1, homog=True):
    """Oracle twin of product_cases.FiniteSizeSolver."""
    from oracle.forms import Problem
    import product_cases as pc

    def bulk(n):
        return {"A": lambda x: np.cos(x[..., 0]) + np.sin(x[..., 1]) + 3.0,
                "B": lambda x: np.cos(x[..., 0]) - np.sin(x[..., 1]) + 3.0,
                "N": lambda x: np.sin(x[..., 0]) * np.cos(x[..., 1]) + 2.0}[n]
    conc_params = [{"name": n, "diffusion coefficient": D, "z": z, "solvated diameter": a, "bulk": bulk(n)}
                   for n, D, z, a in pc.FS_SPECIES]
    flow = ["migration", "poisson"]
    flow += ["diffusion", "finite size"] if kind == "finite size" else ["diffusion finite size_" + kind]
    physical_params = {"flow": flow, "F": 1.0, "R": 1.0, "T": 1.0, "vacuum permittivity": 2.0,
                       "relative permittivity": 1e-1, "U_app": Uex, "Avogadro constant": 0.1}
    markers = {"bulk dirichlet": (1, 2, 3, 4), "applied": (1, 2, 3, 4)}
    return Problem(conc_params, physical_params, markers, family="CG", p=p,
                   homog_params=[dict(r) for r in pc.FS_HOMOG] if homog else [])


def gupta_problem(SUPG=False):
    """Oracle twin of product_cases.GuptaSolver (examples/gupta_advection_migration.py)."""
    from oracle.forms import Problem
    import product_cases as pc
    Lx = pc.GUPTA_LX
    conc_params = [{"name": n, "diffusion coefficient": D, "bulk": b, "z": z} for n, D, b, z in pc.GUPTA_SPECIES]
    physical_params = {"flow": ["diffusion", "electroneutrality", "migration", "advection"],
                       "F": 96485.3329, "R": 8.3144598, "T": 273.15 + 25.}

    def current(cef):
        def curr(u, x):
            X = x[..., 0]
            return cef * 50. * ((X >= 1e-3) & (X < Lx - 1e-3)).astype(float)
        return curr
    echem_params = [{"reaction": current(c), "stoichiometry": dict(st), "electrons": ne, "boundary": "electrode"}
                    for c, st, ne in pc.GUPTA_CURRENTS]
    markers = {"inlet": (1,), "outlet": (2,), "bulk": (4,), "electrode": (3,)}

    def vel(x):
        v = np.zeros(x.shape, dtype=x.dtype)
        v[..., 0] = 1.91 * x[..., 1]
        return v
    return Problem(conc_params, physical_params, markers, echem_params=echem_params, family="CG", vel=vel, SUPG=SUPG,
                   homog_params=[dict(r) for r in pc.GUPTA_HOMOG])


def gupta_oracle_mesh(nx=6, ny=4, n_layer=3):
    from oracle.mesh import tensor_mesh
    from echemfem_b200.meshes import RectangleBoundaryLayerMesh
    import product_cases as pc
    m = RectangleBoundaryLayerMesh(nx, ny, pc.GUPTA_LX, pc.GUPTA_LY, n_layer, 1e-6, boundary=(3,))
    return tensor_mesh([m.axes[0], m.axes[1]])


def catalyst_layer_problem(p=1):
    """Oracle twin of product_cases.CatalystLayerSolver."""
    from oracle.forms import Problem
    import product_cases as pc
    P, K = pc.CL_PARAMS, pc.CL_RATES
    epsS = P["porosity"] * P["saturation"]

    def bulk_reaction(y, x):
        CO2, OH, H, CO3, HCO3 = y[0], y[1], y[2], y[3], y[4]
        r1 = K["k1f"] * CO2 * OH - K["k1b"] * HCO3
        r2 = K["k2f"] * HCO3 * OH - K["k2b"] * CO3
        rw = K["kwf"] - K["kwb"] * H * OH
        return [epsS * (-r1), epsS * (-r1 - r2 + rw), epsS * rw, epsS * r2, epsS * (r1 - r2), 0.0]

    def tafel(i0, alpha, U0):
        def r(u, x):
            eta = u[7] - u[6] - (U0 + np.log(u[2] / 0.2))
            return -i0 * np.exp(-alpha * eta)
        return r
    conc_params = []
    for n, D, z, b, kmt, gas in pc.CL_SPECIES:
        cp = {"name": n, "diffusion coefficient": D, "z": z, "bulk": b, "mass transfer coefficient": kmt}
        if gas is not None:
            cp["gas"] = gas
        conc_params.append(cp)
    physical_params = dict(P, flow=["diffusion", "migration", "electroneutrality full", "porous"], U_app=-0.3)
    physical_params["bulk reaction"] = bulk_reaction
    echem_params = [{"reaction": tafel(i0, al, U0), "electrons": ne, "stoichiometry": dict(nu)}
                    for i0, al, U0, nu, ne in pc.CL_ECHEM]
    markers = {"applied": (2,), "gas": (2,), "bulk": (1,)}
    return Problem(conc_params, physical_params, markers, echem_params=echem_params, family="CG", p=p)


def bmcsl_1d_problem(p=2, sigma=-0.3):
    """Oracle twin of product_cases.BMCSL1DSolver."""
    from oracle.forms import Problem
    conc_params = [{"name": "Cl", "diffusion coefficient": 2.0, "bulk": 1.0, "z": -1, "solvated diameter": 0.0},
                   {"name": "Li", "diffusion coefficient": 1.0, "bulk": 0.9, "z": 1, "solvated diameter": 0.76},
                   {"name": "Cs", "diffusion coefficient": 2.0, "bulk": 0.1, "z": 1, "solvated diameter": 0.66}]
    physical_params = {"flow": ["migration", "poisson", "diffusion finite size_BMCSL"], "F": 1.0, "R": 1.0, "T": 1.0,
                       "U_app": 0.0, "vacuum permittivity": 0.05, "relative permittivity": 1.0,
                       "Avogadro constant": 0.2, "surface charge density": sigma}
    markers = {"bulk dirichlet": (1,), "applied": (1,), "poisson neumann": (2,)}
    return Problem(conc_params, physical_params, markers, family="CG", p=p)


def bmcsl_1d_oracle_mesh():
    from oracle.mesh import tensor_mesh
    from echemfem_b200.meshes import IntervalBoundaryLayerMesh
    return tensor_mesh([IntervalBoundaryLayerMesh(6, 1.0, 4, 0.1, boundary=(2,)).axes[0]])
