"""`EchemSolver` -- the reference's subclass contract on top of the B200 Newton step.

Same constructor, hooks, attributes and methods as the reference class
(echemfem/solver.py:33-71, 528, 722, 927, 1276, 2586-2626), so example and test
scripts drive it unchanged.  Where the reference hands a UFL form to
`NonlinearVariationalSolver` (solver.py:714-718) this class hands the same
weak form -- restated below with `echemfem_b200.symbolic` -- to
`echemfem_b200.compiler`, which differentiates it and emits the pointwise
integrands the hand-written CUDA kernels call; `solve()` then runs the Newton
loop on the device (`echemfem_b200.newton`).

The weak form is assembled from four term families instead of being written
out per boundary: `_interior_penalty` (SIP), `_weak_dirichlet` (Nitsche),
`_upwind` and pointwise sources.  Each cites the reference lines it restates.
"""
from . import symbolic as S
from .symbolic import (Constant, Expr, inner, dot, grad, avg, jump, conditional,
                       max_value, sqrt, ln, as_expr, as_vector)
from .function import Function, FunctionSpace, MixedFunctionSpace

_FLOW_KEYS = ("advection", "diffusion", "migration", "finite size", "electroneutrality",
              "electroneutrality full", "poisson", "porous", "darcy", "advection gas only",
              "diffusion finite size_BK", "diffusion finite size_CS", "diffusion finite size_SP",
              "diffusion finite size_SMPNP", "diffusion finite size_BMCSL")


class _Comm:
    """Single-process stand-in for COMM_WORLD (solver.py:84)."""
    size = 1
    rank = 0

    def allreduce(self, x, op=None):
        return x


def _is_number(a):
    return isinstance(a, (int, float))


class EchemSolver:
    # ------------------------------------------------------------------ hooks
    def set_boundary_markers(self):
        raise NotImplementedError('method needs to be implemented by solver')

    def set_velocity(self):
        raise NotImplementedError('method needs to be implemented by solver')

    def neumann(self, C, conc_params, u):
        raise NotImplementedError('method needs to be implemented by solver')

    def poisson_neumann(self, u):
        raise NotImplementedError('method needs to be implemented by solver')

    # ------------------------------------------------------------ construction
    def __init__(self, conc_params, physical_params, mesh, echem_params=[], homog_params=[],
                 gas_params=[], stats_file="", overwrite_stats_file=False, p=1, family="DG",
                 p_penalty=None, SUPG=False, cylindrical=False):
        self.physical_params = physical_params
        self.conc_params = conc_params
        self.echem_params = echem_params
        self.homog_params = homog_params
        self.gas_params = gas_params
        self.num_c = len(conc_params)
        self.num_g = len(gas_params)
        self.stats_file = stats_file
        self.overwrite_stats_file = overwrite_stats_file
        self.comm = _Comm()
        self.mesh = mesh
        S.set_geometric_dimension(mesh.dim)

        self.set_boundary_markers()
        for name, ids in self.boundary_markers.items():
            if isinstance(ids, int):
                print("*** WARNING: boundary marker converted from int to tuple")
                self.boundary_markers[name] = (ids,)

        self.flow = {k: False for k in _FLOW_KEYS}
        self.SUPG = SUPG
        self.save_solutions = True
        for physics in physical_params["flow"]:
            self.flow[physics] = True
        self._validate(family)
        if gas_params or self.flow["darcy"] or self.flow["advection gas only"]:
            raise NotImplementedError("gas phase / Darcy models are outside the accelerated path")

        # kept species; the eliminated one is the last unless tagged (solver.py:144-162)
        self.idx_c = list(range(self.num_c))
        if self.flow["electroneutrality"]:
            tagged = [k for k, c in enumerate(conc_params) if c.get("eliminated") is not None]
            assert len(tagged) < 2, "Can only eliminate one aqueous concentration"
            if not tagged:
                conc_params[-1]["eliminated"] = True
                tagged = [self.num_c - 1]
            self.i_el = tagged[0]
            self.idx_c.remove(self.i_el)
            print("Eliminating species: " + conc_params[self.i_el]["name"])
            assert not conc_params[self.i_el]["z"] == 0, \
                "Eliminated species must have non-zero charge. Add \"eliminated\": True to a species with non-zero charge."
        self.num_liquid = len(self.idx_c)
        self.num_gas = 0
        self.num_mass = self.num_liquid

        self.poly_degree = self.poly_degreeU = self.poly_degreep = p
        self.family = family
        self.C_IP = Constant(10.)
        if p_penalty is None:
            p_penalty = Constant(p)
        self.penalty_degree = self.penalty_degreeU = self.penalty_degreep = p_penalty

        # measures: degree p+1 everywhere (solver.py:214-251)
        weight = None
        if cylindrical:
            print("Using cylindrical coordinates")
            if mesh.topological_dimension() != 2:
                raise NotImplementedError('Cylindrical coordinates require a 2D mesh')
            weight = S.coordinate(2)[0]
        self._quad_degree = p + 1
        self.dx = S.Measure("cell", degree=p + 1, weight=weight)
        self.ds = S.Measure("exterior_facet", degree=p + 1, weight=weight)
        self.dS = S.Measure("interior_facet", degree=p + 1, weight=weight)

        self.set_function_spaces(mesh.layers, family)
        has_U = self.flow["poisson"] or self.flow["electroneutrality"] or self.flow["electroneutrality full"]
        spaces = [self.V] * self.num_mass
        if has_U:
            spaces.append(self.Vu)
            if self.flow["porous"]:
                spaces.append(self.Vu)
        self.W = MixedFunctionSpace(spaces)
        self.vector = False
        self.vector_mix = False

        d = mesh.dim
        nf = len(spaces)
        self.u = Function(self.W, name="solution")
        self.u._symbol_fields = nf
        us = [S.FieldRef("u", j, d) for j in range(nf)]
        v = [S.FieldRef("v", j, d) for j in range(nf)]
        self._us, self._v = us, v
        self._aux = []                           # auxiliary data fields (Functions used as coefficients)

        self.i_c = {conc_params[k]["name"]: i for i, k in enumerate(self.idx_c)}
        self.i_g = {}
        if has_U:
            self.i_Ul = self.num_mass
            if self.flow["porous"]:
                self.i_Us = self.num_mass + 1

        self.Sw = physical_params.get("saturation")
        if self.Sw is None:
            self.Sw = Constant(1.0)

        # applied potential: number -> Constant, expression -> interpolated data field (solver.py:373-380)
        U_app = physical_params.get("U_app")
        if U_app is not None:
            if isinstance(U_app, (Constant, Function)):
                self.U_app = U_app
            elif _is_number(U_app):
                self.U_app = Constant(U_app)
            else:
                self.U_app = Function(self.Vu).interpolate(U_app)
            if isinstance(self.U_app, Function):
                self.register_coefficient(self.U_app)

        if self.flow["advection"]:
            self.set_velocity()
            self.vel = self._ingest_velocity(self.vel)
        if self.homog_params:
            self.setup_bulk_reactions()

        self.setup_forms(us, v)
        self.init_solver_parameters()
        self._engine = None

    def _ingest_velocity(self, vel):
        """`set_velocity` may leave a symbolic expression in `self.vel` (most examples) or a DISCRETE field, as
        `FlowSolver.vel` is (flow_solver.py:114; examples/simple_flow_battery.py:98-114): a `VectorFunction`, or a
        sequence of scalar `Function`s.  Discrete components become data coefficients of the form (auxiliary nodal
        fields next to U_app), interpolated in the kernels like the unknowns.  They must live on the scalar space
        of the concentrations -- a field of another space is transferred by nodal interpolation with
        `VectorFunction(VectorFunctionSpace(mesh, family, p)).interpolate(...)` first."""
        from .function import VectorFunction
        comps = None
        if isinstance(vel, VectorFunction):
            comps = vel.components
        elif isinstance(vel, (list, tuple)) and any(isinstance(c, Function) for c in vel):
            comps = list(vel)
        if comps is None:
            return vel
        d = self.mesh.dim
        if len(comps) != d:
            raise ValueError("velocity field has %d components on a %dD mesh" % (len(comps), d))
        out = []
        for c in comps:
            if isinstance(c, Function):
                sp_ = c.function_space()
                if (sp_.family, sp_.degree, sp_.num_nodes) != (self.V.family, self.V.degree, self.V.num_nodes):
                    raise NotImplementedError("discrete velocity components must live on the solver's scalar space "
                                              "(%s%d); interpolate the flow field there first" % (self.V.family, self.V.degree))
                self.register_coefficient(c)
                out.append(c._as_expr())
            else:
                out.append(c)
        self.velocity_field = vel
        return as_vector(out)

    def register_coefficient(self, fn):
        """Make a data `Function` usable inside forms (it becomes auxiliary field a{j})."""
        if fn._symbol is None:
            fn._symbol = ("a", len(self._aux))
            self._aux.append(fn)
        return fn

    def _validate(self, family):
        f = self.flow
        if (not f["migration"]) and f["electroneutrality"]:
            raise NotImplementedError('Electroneutrality constraint requires Migration')
        if f["migration"] and not (f["electroneutrality"] or f["poisson"] or f["electroneutrality full"]):
            raise NotImplementedError('Migration requires Electroneutrality or Poisson')
        if f["poisson"] and (f["electroneutrality"] or f["electroneutrality full"]):
            raise NotImplementedError('Electroneutrality not compatible with Poisson')
        for physics in self.physical_params["flow"]:
            if "finite size" in physics and family == "DG":
                raise NotImplementedError('finite size effects only implemented with CG')
        if f["darcy"] and (not f["porous"]):
            raise NotImplementedError('Darcy requires a porous model')
        if self.gas_params and not f["darcy"]:
            raise NotImplementedError('Two-phase currently only supported for Darcy flow')
        if f["darcy"] and not (f["advection"] or f["advection gas only"]):
            raise NotImplementedError('Darcy requires advection or advection gas only')
        if not self.gas_params and f["advection gas only"]:
            raise NotImplementedError('advection gas only requires gaseous species')

    def set_function_spaces(self, layers, family):
        # DQ x DG on extruded meshes is the same tensor DG space here (solver.py:1434-1459)
        self.V = FunctionSpace(self.mesh, family, self.poly_degree)
        self.Vu = FunctionSpace(self.mesh, family, self.poly_degreeU)
        self.Vp = FunctionSpace(self.mesh, family, self.poly_degreep)

    # ------------------------------------------------------------------ forms
    def setup_forms(self, us, v):
        self.Form, self.bcs = self.steady_forms(us, v)

    def steady_forms(self, us, v):
        """Sum of all equations (solver.py:401-488)."""
        cps = self.conc_params
        pp = self.physical_params
        F = S.Form()
        bcs = []
        for i, k in enumerate(self.idx_c):
            shifted = self.flow["electroneutrality full"] and i == self.num_liquid - 1
            w = v[i + 1] if shifted else v[i]
            cps[k]["i_c"] = i
            a, bc = self.mass_conservation_form(us[i], w, cps[k], u=us)
            F += cps[k].get("residual weight", 1.0) * a
            bcs += bc
        if self.flow["electroneutrality"]:
            a, bc = self.charge_conservation_form(us, v[self.num_mass], cps)
            F += a
            bcs += bc
        elif self.flow["poisson"]:
            a, bc = self.potential_poisson_form(us, v[self.num_mass], cps, solid=False)
            F += a
            bcs += bc
        elif self.flow["electroneutrality full"]:
            a, bc = self.electroneutrality_form(us, v[self.num_liquid - 1], cps)
            F += pp.get("electroneutrality residual weight", 1.0) * a
            bcs += bc
        if self.flow["porous"] and (self.flow["poisson"] or self.flow["electroneutrality"]
                                    or self.flow["electroneutrality full"]):
            a, bc = self.potential_poisson_form(us, v[self.num_mass + 1], cps, solid=True)
            F += pp.get("solid potential residual weight", 1.0) * a
            bcs += bc
        return F, bcs

    # -- term families -----------------------------------------------------------
    def _penalty(self, degree):
        """C_IP max(p^2, 1) / he with he = CellVolume / FacetArea (solver.py:1491-1492)."""
        return self.C_IP * max_value(degree ** 2, 1) * S.facet_area() / S.cell_volume()

    def _interior_penalty(self, kappa, w, test, sigma=None):
        """SIP for  -div(kappa grad w)  (solver.py:1518-1520, 1927-1930, 1981-1983)."""
        n = S.facet_normal(self.mesh.dim)
        a = -inner(avg(kappa * grad(w)), jump(test, n)) * self.dS
        a -= inner(avg(kappa * grad(test)), jump(w, n)) * self.dS
        if sigma is not None:
            a += avg(sigma * kappa) * jump(w) * jump(test) * self.dS
        return a

    def _weak_dirichlet(self, kappa, w, w0, test, ids, sigma=None):
        """Nitsche boundary terms for  w = w0  (solver.py:1525-1529, 2003-2006)."""
        n = S.facet_normal(self.mesh.dim)
        a = inner(kappa * (w0 - w) * n, grad(test)) * self.ds(ids)
        a -= inner(kappa * grad(w), n * test) * self.ds(ids)
        if sigma is not None:
            a += inner(sigma * kappa * (w - w0) * n, test * n) * self.ds(ids)
        return a

    def _upwind(self, flow, C, test):
        """(v+ - v-)(flow_N+ C+ - flow_N- C-),  flow_N = (flow.n + |flow.n|)/2  (solver.py:1652-1654)."""
        n = S.facet_normal(self.mesh.dim)
        fn = dot(flow, n)
        flow_N = S.positive_part(fn)         # 0.5 * (fn + abs(fn)), kept as one node (symbolic.PosPart)
        return (test('+') - test('-')) * (flow_N('+') * C('+') - flow_N('-') * C('-')) * self.dS

    def _packing(self, u, powers=(3,)):
        """sum_i (pi/6)^0 N_A a_i^m c_i over species with a solvated diameter."""
        NA = self.physical_params["Avogadro constant"]
        out = [0.0] * len(powers)
        for i in range(self.num_c):
            ai = self.conc_params[i].get("solvated diameter")
            if ai is not None and ai != 0.0:
                for m, pw in enumerate(powers):
                    out[m] = out[m] + NA * ai ** pw * u[i]
        return out

    # -- species transport -------------------------------------------------------------
    def mass_conservation_form(self, C, test_fn, conc_params, u=None):
        """Transport equation of one kept species (solver.py:1461-1733)."""
        pp = self.physical_params
        bm = self.boundary_markers
        DG = self.family == "DG"
        i_c = conc_params["i_c"]
        D = conc_params.get("diffusion coefficient")
        C_0 = conc_params.get("bulk")
        C_gas = conc_params.get("gas")
        if self.flow["porous"]:
            D = self.effective_diffusion(D)
        n = S.facet_normal(self.mesh.dim)
        sigma = self._penalty(self.penalty_degree) if DG else None
        a = S.Form()
        bcs = []

        source = None
        if pp.get("bulk reaction") is not None:
            source = pp["bulk reaction"](u)[i_c]
            if source != 0:
                a -= source * test_fn * self.dx                                # :1497-1500
            else:
                source = None

        charged = False
        if self.flow["poisson"] or self.flow["electroneutrality"] or self.flow["electroneutrality full"]:
            U = u[self.i_Ul]
        if self.flow["migration"]:
            z = conc_params["z"]
            charged = z != 0.0
            K = D * z * pp["F"] / pp["R"] / pp["T"]                            # :1509
        Far = pp.get("F")

        if self.flow["diffusion"]:
            a += inner(D * grad(C), grad(test_fn)) * self.dx                   # :1515
            if DG:
                a += self._interior_penalty(D, C, test_fn, sigma)
            for key, target in (("bulk dirichlet", C_0), ("gas", C_gas)):
                ids = bm.get(key)
                if ids is None or (key == "gas" and target is None):
                    continue
                if DG:
                    a += self._weak_dirichlet(D, C, target, test_fn, ids, sigma)
                bcs.append(("dirichlet", i_c, target, ids))                    # :1531, 1540
            if self.flow["finite size"]:                                       # :1543-1552
                NA = pp["Avogadro constant"]
                den, num = 1.0, 0.0
                for i in range(self.num_c):
                    ai = self.conc_params[i].get("solvated diameter")
                    if ai is not None and ai != 0.0:
                        den = den - NA * ai ** 3 * u[i]
                        num = num + NA * ai ** 3 * grad(u[i])
                a += D * C * inner(num / den, grad(test_fn)) * self.dx

        mu_ex = self._excess_potential(conc_params, u)
        if mu_ex is not None:                                                  # :1554-1636
            a += D * C * inner(grad(ln(C) + mu_ex), grad(test_fn)) * self.dx
            if bm.get("bulk dirichlet") is not None:
                bcs.append(("dirichlet", i_c, C_0, bm["bulk dirichlet"]))

        flow = None
        if self.flow["advection"] and charged:
            flow = self.vel - K * grad(U)                                      # :1644
        elif charged:
            flow = -K * grad(U)
        elif self.flow["advection"]:
            flow = self.vel
        if flow is not None:
            a -= inner(C * flow, grad(test_fn)) * self.dx                      # :1649
            if DG:
                a += self._upwind(flow, C, test_fn)
            elif self.SUPG:                                                    # :1655-1668
                hk = sqrt(2) * S.cell_volume() / S.cell_diameter()
                u_norm = inner(flow, flow) ** 0.5
                Pe = hk / 2. * u_norm / D
                Pe_f = conditional(Pe > 1, 1. - 1. / Pe, 0)
                delta_k = hk / 2. / u_norm * Pe_f
                a += delta_k * inner(dot(flow, grad(C)), dot(flow, grad(test_fn))) * self.dx
                if source is not None:
                    a -= source * delta_k * dot(flow, grad(test_fn)) * self.dx
            applied = bm.get("applied")
            liquid_applied = bm.get("liquid applied")
            idx_app = None
            if (applied is not None and not self.flow["porous"]) or liquid_applied is not None:
                idx_app = liquid_applied if self.flow["porous"] else applied
            if idx_app is not None:                                            # :1681-1685
                fn = dot(flow, n)
                a += conditional(fn < 0, test_fn * fn * C_0, 0.0) * self.ds(idx_app)
                a += conditional(fn > 0, test_fn * fn * C, 0.0) * self.ds(idx_app)
            for key in ("inlet", "outlet"):                                    # :1687-1700
                ids = bm.get(key)
                if ids is None:
                    continue
                if idx_app is not None:
                    ids = tuple(set(ids) - set(idx_app))
                vn = dot(self.vel, n)
                if key == "inlet":
                    if C_0 != 0.0:
                        a += conditional(vn < 0, test_fn * vn * C_0, 0.0) * self.ds(ids)
                else:
                    a += conditional(vn > 0, test_fn * vn * C, 0.0) * self.ds(ids)

        if bm.get("bulk") is not None and conc_params.get("mass transfer coefficient") is not None:
            a -= conc_params["mass transfer coefficient"] * (C_0 - C) * test_fn * self.ds(bm["bulk"])

        name = conc_params["name"]                                             # :1708-1726
        for echem in self.echem_params:
            if name not in echem["stoichiometry"]:
                continue
            rate = echem["stoichiometry"][name] / echem["electrons"] / Far * echem["reaction"](u)
            if not self.flow["porous"]:
                a -= test_fn * rate * self.ds(bm.get(echem["boundary"]))
            else:
                av = pp["specific surface area"] * self.Sw
                a -= av * test_fn * rate * self.dx

        if bm.get("neumann") is not None:                                      # :1728-1731
            a -= test_fn * self.neumann(C, conc_params, u) * self.ds(bm["neumann"])
        return a, bcs

    def _excess_potential(self, conc_params, u):
        """mu_ex of the 'diffusion finite size_*' closures (solver.py:1554-1633)."""
        f = self.flow
        pi = S.pi
        if f["diffusion finite size_BK"]:
            phi, = self._packing(u)
            return -ln(1 - phi)
        if f["diffusion finite size_CS"]:
            phi, = self._packing(u)
            return phi * (8 - 9 * phi + 3 * phi ** 2) / (1 - phi) ** 3
        if f["diffusion finite size_SP"]:
            phi, = self._packing(u)
            return -ln(1 - 8 * phi)
        if f["diffusion finite size_SMPNP"]:
            phi, = self._packing(u)
            a0 = 2.3e-10
            betaj = (conc_params.get("solvated diameter") / a0) ** 3
            return -betaj * ln(1 - phi)
        if f["diffusion finite size_BMCSL"]:
            p0, p1, p2, p3 = [(pi / 6) * t for t in self._packing(u, (0, 1, 2, 3))]
            ai = conc_params.get("solvated diameter")
            return (-(1 + (2 * p2 ** 3 * ai ** 3 / (p3 ** 3)) - (3 * p2 ** 2 * ai ** 2 / (p3 ** 2))) * ln(1 - p3)
                    + (3 * p2 * ai + 3 * p1 * ai ** 2 + p0 * ai ** 3) / (1 - p3)
                    + (6 * p1 * p2 * ai ** 3 + 9 * p2 ** 2 * ai ** 2) / (2 * (1 - p3) ** 2)
                    + (3 * p2 ** 3 * ai ** 3) / (1 - p3) ** 3
                    + ((3 * p2 ** 2 * ai ** 2) / p3) * ((1 - 3 * p3 * 0.5) / (1 - p3) ** 2)
                    - (p2 ** 3 * ai ** 3) * (4 * p3 ** 2 - 5 * p3 + 2) / (p3 ** 2 * (1 - p3) ** 3))
        return None

    # -- charge conservation (electroneutral) -------------------------------------------
    def charge_conservation_form(self, u, test_fn, conc_params, W=None, i_bc=None):
        """Liquid-potential equation with one species eliminated (solver.py:1866-2013)."""
        pp = self.physical_params
        bm = self.boundary_markers
        Far, R, T = pp["F"], pp["R"], pp["T"]
        DG = self.family == "DG"
        sigmaU = self._penalty(self.penalty_degreeU) if DG else None
        U = u[self.i_Ul]
        a = S.Form()
        bcs = []
        K_U = 0.0
        neutral_sum = 0.0
        reactions = pp["bulk reaction"](u) if pp.get("bulk reaction") is not None else None
        if self.flow["porous"]:
            av = pp["specific surface area"] * self.Sw

        for k in self.idx_c + [self.i_el]:
            cp = conc_params[k]
            z = cp["z"]
            if z == 0.0:
                continue
            C_0 = cp.get("bulk")
            C_ND = cp.get("C_ND", 1.0)
            if C_ND is None:
                C_ND = 1.0
            C_gas = cp.get("gas")
            if k != self.i_el:
                C = u[cp["i_c"]]
                neutral_sum = neutral_sum + z * C * C_ND                       # :1914
            else:
                C = -neutral_sum / z / C_ND                                    # :1916
            D = cp["diffusion coefficient"]
            if self.flow["porous"]:
                D = self.effective_diffusion(D)
            K_U = K_U + z ** 2 * Far ** 2 * D * C / R / T * C_ND               # :1920
            zDF = z * D * Far * C_ND
            if self.flow["diffusion"]:
                a += inner(zDF * grad(C), grad(test_fn)) * self.dx             # :1924
                if DG:
                    a += self._interior_penalty(zDF, C, test_fn)               # no penalty term (:1931)
                    if bm.get("bulk dirichlet") is not None:
                        a += self._weak_dirichlet(zDF, C, C_0, test_fn, bm["bulk dirichlet"])
                    if bm.get("gas") is not None and C_gas is not None:
                        a += self._weak_dirichlet(zDF, C, C_gas, test_fn, bm["gas"])
            if reactions is not None and reactions[k] != 0.0:
                a -= C_ND * z * Far * reactions[k] * test_fn * self.dx        # :1946
            if bm.get("bulk") is not None and cp.get("mass transfer coefficient") is not None:
                a -= C_ND * z * Far * cp["mass transfer coefficient"] * (C_0 - C) * test_fn * self.ds(bm["bulk"])
            for echem in self.echem_params:                                    # :1952-1969
                if cp["name"] not in echem["stoichiometry"]:
                    continue
                rate = z * echem["stoichiometry"][cp["name"]] / echem["electrons"] * echem["reaction"](u)
                if not self.flow["porous"]:
                    a -= C_ND * test_fn * rate * self.ds(bm.get(echem["boundary"]))
                else:
                    a -= C_ND * av * test_fn * rate * self.dx
            if bm.get("neumann") is not None:
                a -= C_ND * z * Far * test_fn * self.neumann(C, cp, u) * self.ds(bm["neumann"])

        a += inner(K_U * grad(U), grad(test_fn)) * self.dx                     # :1978
        if DG:
            a += self._interior_penalty(K_U, U, test_fn, sigmaU)

        targets = []
        if bm.get("bulk") is not None:
            targets.append((bm["bulk"], Constant(0)))
        if bm.get("applied") is not None and not self.flow["porous"]:
            targets.append((bm["applied"], self.U_app))
        if bm.get("liquid applied") is not None:
            targets.append((bm["liquid applied"], self.U_app))
        for ids, U_0 in targets:
            if DG:
                a += self._weak_dirichlet(K_U, U, U_0, test_fn, ids, sigmaU)
            bcs.append(("dirichlet", self.i_Ul if i_bc is None else i_bc, U_0, ids))
        return a, bcs

    def electroneutrality_form(self, u, test_fn, conc_params):
        """Algebraic constraint sum z_k c_k = 0, CG only (solver.py:2015-2046)."""
        bm = self.boundary_markers
        a = S.Form()
        for i in range(self.num_c):
            z = conc_params[i].get("z")
            if z != 0:
                a += z * u[i] * test_fn * self.dx
        bcs = []
        applied = bm.get("applied")
        is_applied = (applied is not None and not self.flow["porous"]) or bm.get("liquid applied") is not None
        if is_applied:
            bcs.append(("dirichlet", self.num_mass, self.U_app, applied))
        elif bm.get("bulk") is not None:
            bcs.append(("dirichlet", self.num_mass, Constant(0), bm["bulk"]))
        return a, bcs

    # -- Poisson equations ---------------------------------------------------------------
    def potential_poisson_form(self, u, test_fn, conc_params, solid=False, W=None, i_bc=None):
        """Liquid or solid potential equation (solver.py:2048-2234)."""
        pp = self.physical_params
        bm = self.boundary_markers
        Far, R, T = pp["F"], pp["R"], pp["T"]
        eps_0 = pp.get("vacuum permittivity")
        eps_r = pp.get("relative permittivity")
        DG = self.family == "DG"
        sigmaU = self._penalty(self.penalty_degreeU) if DG else None
        porous = self.flow["porous"]
        i_u = self.i_Us if (porous and solid) else self.i_Ul
        U = u[i_u]
        a = S.Form()
        bcs = []
        if porous:
            av = pp["specific surface area"] * self.Sw
        if solid:
            K_U = self.effective_diffusion(pp["solid conductivity"], phase="solid")
        elif eps_r is not None and eps_0 is not None:
            K_U = eps_r * eps_0
        elif pp.get("liquid conductivity"):
            K_U = pp["liquid conductivity"]
            if porous:
                K_U = self.effective_diffusion(K_U, phase="liquid")
        else:
            K_U = 0.0
        reactions = pp["bulk reaction"](u) if pp.get("bulk reaction") is not None else None

        if not solid:
            if self.flow["electroneutrality"]:
                order, i_el = self.idx_c + [self.i_el], self.i_el
            else:
                order, i_el = list(range(self.num_liquid)), None
            neutral_sum = 0.0
            own_conductivity = (eps_r is None or eps_0 is None) and pp.get("liquid conductivity") is None
            for k in order:
                cp = conc_params[k]
                z = cp.get("z")
                if z is None:
                    continue
                if k != i_el:
                    C = u[cp["i_c"]]
                    neutral_sum = neutral_sum + z * C
                else:
                    C = -neutral_sum / z
                if reactions is not None and self.flow["electroneutrality"] and reactions[k] != 0.0:
                    a -= z * Far * reactions[k] * test_fn * self.dx            # :2122-2126
                if bm.get("neumann") is not None and z != 0.0 and self.flow["electroneutrality"]:
                    a -= z * Far * test_fn * self.neumann(C, cp, u) * self.ds(bm["neumann"])
                if z == 0:
                    continue
                D = cp["diffusion coefficient"]
                if porous:
                    D = self.effective_diffusion(D)
                if own_conductivity:
                    K_U = K_U + z ** 2 * Far ** 2 * D * C / R / T              # :2138
                elif eps_r:
                    a -= Far * z * C * test_fn * self.dx                       # :2141
                if not self.flow["poisson"]:                                   # :2144-2161
                    for echem in self.echem_params:
                        if cp["name"] not in echem["stoichiometry"]:
                            continue
                        rate = z * echem["stoichiometry"][cp["name"]] / echem["electrons"] * echem["reaction"](u)
                        if not porous:
                            a -= test_fn * rate * self.ds(bm.get(echem["boundary"]))
                        else:
                            a -= av * test_fn * rate * self.dx
        if solid:
            for echem in self.echem_params:
                a += av * test_fn * echem["reaction"](u) * self.dx             # :2162-2164
        elif self.flow["poisson"]:
            for echem in self.echem_params:
                a -= av * test_fn * echem["reaction"](u) * self.dx             # :2165-2167 (porous only)

        a += inner(K_U * grad(U), grad(test_fn)) * self.dx                     # :2171
        if DG:
            a += self._interior_penalty(K_U, U, test_fn, sigmaU)

        if reactions is not None:                                              # :2178-2186
            n_r = self.num_liquid + 1 if solid else self.num_liquid
            if len(reactions) > n_r and reactions[n_r] != 0.0:
                print("Reactions in solid Poisson. for test only")
                a -= reactions[n_r] * test_fn * self.dx

        is_inlet = bm.get("inlet") is not None and solid
        is_bulk = bm.get("bulk") is not None and not solid
        is_applied = bm.get("applied") is not None
        if porous and not solid:
            is_applied = bm.get("liquid applied") is not None
        if is_inlet or is_bulk or is_applied:                                  # :2198-2217
            if is_applied:
                U_0, ids = self.U_app, bm.get("applied")
            elif is_inlet:
                U_0, ids = self.U_app, bm["inlet"]
            else:
                U_0, ids = Constant(0), bm["bulk"]
            if DG:
                a += self._weak_dirichlet(K_U, U, U_0, test_fn, ids, sigmaU)
            bcs.append(("dirichlet", i_u if i_bc is None else i_bc, U_0, ids))

        I_app = pp.get("applied current density")
        if bm.get("applied solid current") and solid:
            a -= I_app * test_fn * self.ds(bm["applied solid current"])
        if bm.get("applied liquid current") and not solid:
            a -= I_app * test_fn * self.ds(bm["applied liquid current"])
        if bm.get("robin") is not None:
            a -= pp["gap capacitance"] * (self.U_app - U) * test_fn * self.ds(bm["robin"])
        if bm.get("poisson neumann") is not None:
            a -= pp["surface charge density"] * test_fn * self.ds(bm["poisson neumann"])
        return a, bcs

    # ------------------------------------------------------------- closures
    def effective_diffusion(self, D, phase="liquid"):
        """Bruggeman correlation (solver.py:1276-1286)."""
        porosity = self.physical_params["porosity"]
        frac = {"solid": 1 - porosity, "liquid": porosity * self.Sw, "gas": porosity * (1 - self.Sw)}[phase]
        return D * frac ** 1.5

    def setup_bulk_reactions(self):
        """Law of mass action from homog_params (solver.py:1222-1274)."""
        def bulk_reaction(u):
            rates = []
            for rxn in self.homog_params:
                kf = rxn["forward rate constant"]
                kb = 0
                if rxn.get("backward rate constant"):
                    kb = rxn["backward rate constant"]
                elif rxn.get("equilibrium constant"):
                    kb = rxn["forward rate constant"] / rxn["equilibrium constant"]
                c_ref = rxn.get("reference concentration") or 1.0
                for name, s in rxn["stoichiometry"].items():
                    order = abs(s)
                    if rxn.get("reaction order") and rxn["reaction order"].get(name):
                        order = rxn["reaction order"][name]
                    act = u[self.i_c[name]] / c_ref
                    if s < 0:
                        kf = kf * act ** order
                    if s > 0:
                        kb = kb * act ** order
                rates.append(c_ref * (kf - kb))
            out = [0] * self.num_c
            for k in self.idx_c:
                name = self.conc_params[k]["name"]
                for rxn, r in zip(self.homog_params, rates):
                    if name in rxn["stoichiometry"]:
                        out[k] = out[k] + rxn["stoichiometry"][name] * r
            if self.flow["porous"]:
                eps_l = self.physical_params["porosity"] * self.Sw
                out = [eps_l * r for r in out]
            return out
        if self.physical_params.get("bulk reaction"):
            print("WARNING: Overwriting bulk reaction from physical_params with the one in homog_params")
        self.physical_params["bulk reaction"] = bulk_reaction

    # --------------------------------------------------------- solver options
    def init_solver_parameters(self, pc_type="lu", custom_solver={}, custom_potential_solver={}):
        """PETSc-style option dictionaries (solver.py:927-1220); 'lu' and 'ilu' are on the device path."""
        if custom_solver:
            print("*** WARNING: Using custom solver options")
            self.solver_parameters = custom_solver
            self.potential_solver_parameters = custom_potential_solver
            return
        self.potential_solver_parameters = {
            "snes_type": "newtonls", "snes_linesearch_type": "l2", "snes_rtol": 1e-8,
            "snes_max_it": 20, "mat_type": "aij", "ksp_type": "preonly", "pc_type": "lu"}
        opts = {"snes_type": "newtonls", "snes_linesearch_type": "l2", "snes_monitor": None,
                "snes_converged_reason": None, "snes_rtol": 1e-16, "snes_atol": 1e-16,
                "snes_divergence_tolerance": 1e6, "snes_max_it": 50}
        if pc_type == "lu":
            opts.update({"mat_type": "aij", "ksp_type": "preonly", "pc_type": "lu",
                         "pc_factor_mat_solver_type": "mumps"})
        elif pc_type == "ilu":
            opts.update({"mat_type": "aij", "ksp_type": "fgmres", "ksp_rtol": 1e-10,
                         "ksp_converged_reason": None, "ksp_monitor": None, "ksp_gmres_restart": 1000,
                         "pc_type": "bjacobi", "sub_pc_type": "ilu", "sub_pc_factor_levels": 0})
        elif pc_type == "gmg":                                                   # solver.py:1190-1201
            opts.update({"mat_type": "aij", "ksp_type": "fgmres", "ksp_gmres_restart": 200,
                         "ksp_converged_reason": None, "ksp_monitor": None, "ksp_rtol": 1e-2, "pc_type": "mg",
                         "mg_coarse_ksp_type": "preonly", "mg_coarse_pc_type": "lu",
                         "mg_levels_ksp_type": "richardson", "mg_levels_ksp_max_it": 1,
                         "mg_levels_pc_type": "bjacobi", "mg_levels_sub_pc_type": "ilu"})
        elif pc_type == "amg" or pc_type.startswith(("block", "cpr")):
            raise NotImplementedError("pc_type %r (multigrid / fieldsplit) is outside the accelerated path; "
                                      "use 'lu' or 'ilu'" % pc_type)
        else:
            raise NotImplementedError('unrecognized preconditioner type')
        self.solver_parameters = opts

    # ----------------------------------------------------------- setup / solve
    def print_solver_info(self):
        print("*** ECHEM")
        print("Activated flow components")
        for component, active in self.flow.items():
            if active:
                print("  " + component)
        print("Function space: %s%d, %d fields, %d dofs" % (self.family, self.poly_degree,
                                                            self.W.num_sub_spaces(), self.W.dim()))

    def setup_solver(self, initial_guess=True, initial_solve=True):
        """Initial guess, potential pre-solve and creation of the device Newton solver (solver.py:528-720)."""
        from .newton import NewtonEngine
        self.print_solver_info()
        u = self.u
        if self._engine is None:
            self._engine = NewtonEngine(self)
        has_U = self.flow["electroneutrality"] or self.flow["poisson"] or self.flow["electroneutrality full"]
        if initial_guess:
            for i, k in enumerate(self.idx_c):                                  # :571-582
                bulk = self.conc_params[k]["bulk"]
                if _is_number(bulk):
                    u.sub(i).assign(bulk)
                else:
                    u.sub(i).interpolate(bulk)
            if self.flow["porous"] and has_U:                                   # :609-643
                u.sub(self.i_Ul).assign(0.0)
                u.sub(self.i_Us).assign(self.U_app)
                if initial_solve:
                    self._engine.solve(frozen_fields=list(range(self.num_mass)),
                                       parameters=self.potential_solver_parameters)
            elif has_U:                                                         # :645-663
                U0 = u.sub(self.i_Ul)
                U0.assign(0.0)
                if hasattr(self, "U_app"):
                    U0.assign(self.U_app / 2)
                if initial_solve:
                    self._engine.solve(frozen_fields=list(range(self.num_mass)),
                                       parameters=self.potential_solver_parameters)
        self.echem_solver = self._engine

    def solve(self):
        """One nonlinear solve (solver.py:722-792)."""
        self._engine.solve(parameters=self.solver_parameters)
        if self.save_solutions:
            self.output_state(self.u)
        if len(self.stats_file) > 0:
            self._engine.write_stats(self.stats_file, self.overwrite_stats_file)

    def output_state(self, u, prefix="results/", initiate=True, **kwargs):
        """VTK output is outside the accelerated path (solver.py:794-925); fields stay on `self.u`."""
        return None


class TransientEchemSolver(EchemSolver):
    """Backward-Euler time stepping around the same Newton step (solver.py:2629-2713)."""

    def setup_forms(self, us, v):
        print("Using transient form")
        self.dt = Constant(1)
        if not hasattr(self, "time"):
            self.time = Constant(0)
        self.u_old = Function(self.W)
        d = self.mesh.dim
        self._old_fields = []
        for j, sub in enumerate(self.u_old.subfunctions):
            self.register_coefficient(sub)
            self._old_fields.append(S.FieldRef("a", sub._symbol[1], d))
        F, bcs = self.steady_forms(us, v)
        for i, k in enumerate(self.idx_c):
            shifted = self.flow["electroneutrality full"] and i == self.num_liquid - 1
            w = v[i + 1] if shifted else v[i]
            a = (us[i] - self._old_fields[i]) / self.dt * w * S.dx()           # :2693 (default quadrature)
            F += self.conc_params[k].get("residual weight", 1.0) * a
        self.Form, self.bcs = F, bcs

    def solve(self, times):
        for i in range(1, len(times)):
            self.dt.assign(times[i] - times[i - 1])
            self.time.assign(times[i])
            self._engine.solve(parameters=self.solver_parameters)
            self.u_old.assign(self.u)
