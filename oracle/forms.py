"""Pointwise weak-form coefficients of the reference's equations (oracle).

Every function returns, for each equation i of the mixed system, the
coefficient `f0[i]` multiplying the test function v_i and `f1[i]` multiplying
grad v_i in the integrand, so that the local residual is

    R_i[a] = sum_q w_q ( f0_i(q) phi_a(q) + f1_i(q) . grad phi_a(q) ).

This is a restatement, in numpy, of

* `mass_conservation_form`      echemfem/solver.py:1461-1733
* `charge_conservation_form`    echemfem/solver.py:1866-2013
* `potential_poisson_form`      echemfem/solver.py:2048-2234 (liquid, non-porous)
* `steady_forms`                echemfem/solver.py:401-488

All arithmetic is dtype-generic so that `oracle/assemble.py` can obtain the
exact Gateaux derivative (what UFL's `derivative` gives the reference) by
complex-step differentiation; `abs` differentiates to `sign`, `conditional`
to a piecewise derivative, as UFL does.

Closures in the problem description take numpy arrays:
  bulk(x), U_app(x), vel(x)                 x: (..., d)
  bulk_reaction(u, x) -> list (num_c)       u: list of field arrays [c_0.., Phi]
  echem["reaction"](u, x) -> array
"""
import numpy as np


def _cabs(a):
    """|a| whose complex-step derivative is sign(a) (UFL: d|f| = sign(f) df)."""
    return a * np.sign(np.real(a))


def _pos(a):
    """conditional(a > 0, 1, 0) evaluated on the real part."""
    return (np.real(a) > 0).astype(float)


def _neg(a):
    return (np.real(a) < 0).astype(float)


def _dot(a, b):
    return np.sum(a * b, axis=-1)


def _evalf(f, x):
    """Evaluate a float-or-callable coefficient at points x (..., d)."""
    if callable(f):
        return f(x)
    return f * np.ones(x.shape[:-1])


class _GV:
    """Value + spatial gradient of a pointwise function of the fields (forward mode): what UFL's
    `grad(f(u))` expands to, restated so that the oracle needs no symbolic algebra.  `v`: (...,),
    `g`: (..., d); dtype-generic (the outer complex step of the Jacobian passes through)."""

    def __init__(self, v, g):
        self.v, self.g = v, g

    @staticmethod
    def lift(a):
        return a if isinstance(a, _GV) else _GV(a, 0.0)

    def __add__(self, o):
        o = _GV.lift(o)
        return _GV(self.v + o.v, self.g + o.g)
    __radd__ = __add__

    def __neg__(self):
        return _GV(-self.v, -self.g)

    def __sub__(self, o):
        return self + (-_GV.lift(o))

    def __rsub__(self, o):
        return _GV.lift(o) + (-self)

    def __mul__(self, o):
        o = _GV.lift(o)
        return _GV(self.v * o.v, _col(self.v) * o.g + _col(o.v) * self.g)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = _GV.lift(o)
        return _GV(self.v / o.v, (_col(o.v) * self.g - _col(self.v) * o.g) / _col(o.v * o.v))

    def __rtruediv__(self, o):
        return _GV.lift(o) / self

    def __pow__(self, n):                       # integer powers only
        return _GV(self.v ** n, _col(n * self.v ** (n - 1)) * self.g)

    def log(self):
        return _GV(np.log(self.v), self.g / _col(self.v))


def _col(a):
    return a[..., None] if isinstance(a, np.ndarray) and a.ndim else a


class Problem:
    """Physics description consumed by the oracle.

    Mirrors the reference constructor's inputs (solver.py:57-71) with numpy
    closures in place of UFL expressions.
    """

    def __init__(self, conc_params, physical_params, boundary_markers,
                 echem_params=(), p=1, family="DG", vel=None, SUPG=False,
                 neumann=None, variant="spectral", C_IP=10.0, p_penalty=None, homog_params=()):
        self.conc_params = [dict(c) for c in conc_params]
        self.physical_params = dict(physical_params)
        self.boundary_markers = {}
        for k, v in boundary_markers.items():
            self.boundary_markers[k] = (v,) if isinstance(v, int) else tuple(v)
        self.echem_params = list(echem_params)
        self.p = p
        self.family = family
        self.variant = variant
        self.vel = vel
        self.SUPG = SUPG
        self.neumann = neumann
        self.C_IP = C_IP
        self.penalty_degree = p if p_penalty is None else p_penalty
        self.homog_params = list(homog_params)
        flow = {k: False for k in ("advection", "diffusion", "migration",
                                   "electroneutrality", "electroneutrality full", "poisson", "porous",
                                   "finite size",
                                   "diffusion finite size_BK", "diffusion finite size_CS",
                                   "diffusion finite size_SP", "diffusion finite size_SMPNP",
                                   "diffusion finite size_BMCSL")}
        for f in physical_params["flow"]:
            if f not in flow:
                raise NotImplementedError("oracle: flow option %r" % f)
            flow[f] = True
        # solver.py:115-126
        if (not flow["migration"]) and flow["electroneutrality"]:
            raise NotImplementedError('Electroneutrality constraint requires Migration')
        if flow["migration"] and not (flow["electroneutrality"] or flow["poisson"] or flow["electroneutrality full"]):
            raise NotImplementedError('Migration requires Electroneutrality or Poisson')
        if flow["poisson"] and (flow["electroneutrality"] or flow["electroneutrality full"]):
            raise NotImplementedError('Electroneutrality not compatible with Poisson')
        if flow["electroneutrality full"] and family == "DG":
            raise NotImplementedError("oracle: 'electroneutrality full' is a CG-only option (solver.py:2018)")
        self.Sw = self.physical_params.get("saturation", 1.0)             # solver.py:365-367 (no two-phase flow here)
        if family == "DG" and any(v for k, v in flow.items() if "finite size" in k):       # solver.py:139-142
            raise NotImplementedError('finite size effects only implemented with CG')
        self.flow = flow
        self.num_c = len(conc_params)
        self.idx_c = list(range(self.num_c))
        self.i_el = None
        # solver.py:148-162
        if flow["electroneutrality"]:
            tagged = [i for i, c in enumerate(self.conc_params) if c.get("eliminated") is not None]
            assert len(tagged) < 2
            self.i_el = tagged[-1] if tagged else self.num_c - 1
            assert self.conc_params[self.i_el]["z"] != 0
            self.idx_c.remove(self.i_el)
        self.num_liquid = len(self.idx_c)
        self.num_mass = self.num_liquid
        self.has_potential = flow["electroneutrality"] or flow["poisson"] or flow["electroneutrality full"]
        self.num_fields = self.num_mass + (1 if self.has_potential else 0)
        self.i_Ul = self.num_mass if self.has_potential else None
        self.i_Us = None
        if flow["porous"] and self.has_potential:              # solver.py:264-267, 327-330
            self.i_Us = self.num_mass + 1
            self.num_fields += 1
        for i, k in enumerate(self.idx_c):
            self.conc_params[k]["i_c"] = i
        # equation (test function) that receives the mass balance of kept species i: with the explicit
        # electroneutrality constraint the LAST species is tested with the potential's test function and the
        # constraint with the last species' one, so that no diagonal block is zero (solver.py:411-414, 455-457)
        self.row = list(range(self.num_liquid))
        self.row_constraint = None
        if flow["electroneutrality full"]:
            self.row[self.num_liquid - 1] = self.num_liquid
            self.row_constraint = self.num_liquid - 1

    # -- helpers ----------------------------------------------------------
    def penalty(self, geo):
        """D_IP = C_IP * max(p^2, 1) / he, he = CellVolume / FacetArea (solver.py:1491-1492)."""
        return self.C_IP * max(self.penalty_degree ** 2, 1) * geo["facet_area"] / geo["cell_volume"]

    def concentrations(self, u, gu):
        """All num_c concentrations (and gradients), eliminated one included (solver.py:1904-1916)."""
        C = [None] * self.num_c
        gC = [None] * self.num_c
        for i, k in enumerate(self.idx_c):
            C[k] = u[i]
            gC[k] = gu[i]
        if self.i_el is not None:
            cp = self.conc_params
            zel = cp[self.i_el]["z"]
            ndel = cp[self.i_el].get("C_ND", 1.0)
            s = 0.0
            gs = 0.0
            for k in self.idx_c:
                z = cp[k]["z"]
                if z != 0.0:
                    nd = cp[k].get("C_ND", 1.0)
                    s = s + z * nd * C[k]
                    gs = gs + z * nd * gC[k]
            C[self.i_el] = -s / zel / ndel
            gC[self.i_el] = -gs / zel / ndel
        return C, gC

    def reaction_u(self, u):
        """The list the reference hands to closures: split(u) (solver.py:308)."""
        return list(u)

    def effective(self, D, phase="liquid"):
        """Bruggeman correlation (solver.py:1276-1286)."""
        if not self.flow["porous"]:
            return D
        eps = self.physical_params["porosity"]
        frac = {"solid": 1.0 - eps, "liquid": eps * self.Sw}[phase]
        return D * frac ** 1.5

    def Dk(self, k):
        return self.effective(self.conc_params[k]["diffusion coefficient"])

    def K_mig(self, k):
        pp = self.physical_params
        cp = self.conc_params[k]
        return self.Dk(k) * cp["z"] * pp["F"] / pp["R"] / pp["T"]

    def conductivity(self, C):
        """K_U of the charge-conservation equation (solver.py:1920)."""
        pp = self.physical_params
        K_U = 0.0
        for k in range(self.num_c):
            cp = self.conc_params[k]
            z = cp["z"]
            if z != 0.0:
                K_U = K_U + z ** 2 * pp["F"] ** 2 * self.Dk(k) * C[k] / pp["R"] / pp["T"] * cp.get("C_ND", 1.0)
        return K_U

    def solid_KU(self):
        return self.effective(self.physical_params["solid conductivity"], "solid")      # :2082-2084

    def poisson_KU(self, C):
        """K_U of the liquid Poisson equation (solver.py:2085-2093, 2136-2138)."""
        pp = self.physical_params
        eps_0 = pp.get("vacuum permittivity")
        eps_r = pp.get("relative permittivity")
        if eps_r is not None and eps_0 is not None:
            return eps_r * eps_0
        if pp.get("liquid conductivity"):
            return pp["liquid conductivity"]
        K_U = 0.0
        for k in range(self.num_liquid):
            cp = self.conc_params[k]
            z = cp.get("z")
            if z is not None and z != 0:
                K_U = K_U + z ** 2 * pp["F"] ** 2 * self.Dk(k) * C[k] / pp["R"] / pp["T"]
        return K_U

    def flow_vector(self, k, x, gU):
        """`flow` of species k (solver.py:1641-1648); None if the species has no transport term."""
        cp = self.conc_params[k]
        adv = self.flow["advection"]
        mig = self.flow["migration"] and cp["z"] != 0.0
        if not (adv or mig):
            return None
        fl = 0.0
        if adv:
            fl = fl + self.vel(x)
        if mig:
            fl = fl - self.K_mig(k) * gU
        return fl

    def mass_action(self, u):
        """Homogeneous bulk reactions by the law of mass action (`setup_bulk_reactions`, solver.py:1222-1274)."""
        name_to_field = {self.conc_params[k]["name"]: i for i, k in enumerate(self.idx_c)}
        rates = []
        for rxn in self.homog_params:
            kf = rxn["forward rate constant"]
            kb = 0.0
            if rxn.get("backward rate constant"):
                kb = rxn["backward rate constant"]
            elif rxn.get("equilibrium constant"):
                kb = rxn["forward rate constant"] / rxn["equilibrium constant"]
            c_ref = rxn["reference concentration"] if rxn.get("reference concentration") else 1.0
            for name, st in rxn["stoichiometry"].items():
                order = abs(st)
                if rxn.get("reaction order") and rxn["reaction order"].get(name):
                    order = rxn["reaction order"][name]
                act = u[name_to_field[name]] / c_ref
                if st < 0:
                    kf = kf * act ** order
                if st > 0:
                    kb = kb * act ** order
            rates.append(kf - kb)
        out = [0.0] * self.num_c
        for k in self.idx_c:
            name = self.conc_params[k]["name"]
            for rxn, r in zip(self.homog_params, rates):
                if name in rxn["stoichiometry"]:
                    c_ref = rxn["reference concentration"] if rxn.get("reference concentration") else 1.0
                    out[k] = out[k] + rxn["stoichiometry"][name] * c_ref * r
        if self.flow["porous"]:
            out = [self.physical_params["porosity"] * self.Sw * r for r in out]
        return out

    def excess_potential(self, cp, u, gu):
        """mu_ex of the 'diffusion finite size_*' options with its gradient (solver.py:1554-1633).
        The sums run over u[i], i < num_c, as in the reference."""
        f = self.flow
        if not any(f["diffusion finite size_" + t] for t in ("BK", "CS", "SP", "SMPNP", "BMCSL")):
            return None
        NA = self.physical_params["Avogadro constant"]

        def packing(power, scale=1.0):
            tot = _GV(0.0, 0.0)
            for j in range(self.num_c):
                aj = self.conc_params[j].get("solvated diameter")
                if aj is not None and aj != 0.0:
                    tot = tot + _GV(u[j], gu[j]) * (scale * NA * aj ** power)
            return tot
        if f["diffusion finite size_BK"]:
            return -((1.0 - packing(3)).log())
        if f["diffusion finite size_CS"]:
            phi = packing(3)
            return phi * (8.0 - 9.0 * phi + 3.0 * phi ** 2) / (1.0 - phi) ** 3
        if f["diffusion finite size_SP"]:
            return -((1.0 - 8.0 * packing(3)).log())
        if f["diffusion finite size_SMPNP"]:
            betaj = (cp.get("solvated diameter") / 2.3e-10) ** 3
            return -betaj * ((1.0 - packing(3)).log())
        p0, p1, p2, p3 = [packing(m, np.pi / 6.0) for m in range(4)]
        ai = cp.get("solvated diameter")
        return (-(1.0 + (2.0 * p2 ** 3 * ai ** 3 / (p3 ** 3)) - (3.0 * p2 ** 2 * ai ** 2 / (p3 ** 2))) * (1.0 - p3).log()
                + (3.0 * p2 * ai + 3.0 * p1 * ai ** 2 + p0 * ai ** 3) / (1.0 - p3)
                + (6.0 * p1 * p2 * ai ** 3 + 9.0 * p2 ** 2 * ai ** 2) / (2.0 * (1.0 - p3) ** 2)
                + (3.0 * p2 ** 3 * ai ** 3) / (1.0 - p3) ** 3
                + ((3.0 * p2 ** 2 * ai ** 2) / p3) * ((1.0 - 3.0 * p3 * 0.5) / (1.0 - p3) ** 2)
                - (p2 ** 3 * ai ** 3) * (4.0 * p3 ** 2 - 5.0 * p3 + 2.0) / (p3 ** 2 * (1.0 - p3) ** 3))

    # -- cell integrals ---------------------------------------------------
    def cell(self, x, u, gu, geo):
        nf = self.num_fields
        d = x.shape[-1]
        zero = np.zeros(x.shape[:-1], dtype=u[0].dtype)
        f0 = [zero.copy() for _ in range(nf)]
        f1 = [np.zeros(x.shape, dtype=u[0].dtype) for _ in range(nf)]
        # magnitude of pointwise sums that cancel analytically (real part: of the terms themselves, imaginary part: of
        # their complex-step derivatives).  The charge source sum_k z_k F R_k of charge-conserving reactions is zero
        # in exact arithmetic; evaluated term by term, as the reference does (solver.py:1944-1946), it leaves
        # eps * max|z_k F R_k| -- a rounding floor no assembled entry reveals.  Read by the tests' tolerance only.
        self.cancel0 = [zero.real.copy().astype(complex) for _ in range(nf)]
        pp = self.physical_params
        C, gC = self.concentrations(u, gu)
        reactions = None
        if self.homog_params:
            reactions = self.mass_action(self.reaction_u(u))                # overrides "bulk reaction", :1271-1274
        elif pp.get("bulk reaction") is not None:
            reactions = pp["bulk reaction"](self.reaction_u(u), x)
        if self.has_potential:
            U, gU = u[self.i_Ul], gu[self.i_Ul]
        else:
            U, gU = None, None
        for i, k in enumerate(self.idx_c):
            cp = self.conc_params[k]
            D = self.Dk(k)
            r = self.row[i]
            if self.flow["diffusion"] and self.flow["finite size"]:         # :1543-1552
                NA = pp["Avogadro constant"]
                den, num = 1.0, 0.0
                for j in range(self.num_c):
                    aj = self.conc_params[j].get("solvated diameter")
                    if aj is not None and aj != 0.0:
                        den = den - NA * aj ** 3 * u[j]
                        num = num + NA * aj ** 3 * gu[j]
                f1[r] = f1[r] + (D * C[k] / den)[..., None] * num
            mu_ex = self.excess_potential(cp, u, gu)
            if mu_ex is not None:                                           # :1554-1636  D C grad(ln C + mu_ex)
                f1[r] = f1[r] + D * gC[k] + (D * C[k])[..., None] * mu_ex.g
            rterm = None
            if reactions is not None and not _is_zero(reactions[i]):
                # solver.py:1497-1500  (note: indexed by i_c)
                rterm = reactions[i]
                f0[r] = f0[r] - rterm
            if self.flow["diffusion"]:
                f1[r] = f1[r] + D * gC[k]                                   # :1515
            fl = self.flow_vector(k, x, gU)
            if fl is not None:
                f1[r] = f1[r] - C[k][..., None] * fl                        # :1649
                if self.family != "DG" and self.SUPG:                       # :1655-1668
                    hk = np.sqrt(2.0) * geo["cell_volume"] / geo["cell_diameter"]
                    u_norm = _dot(fl, fl) ** 0.5
                    Pe = hk / 2.0 * u_norm / D
                    Pe_f = np.where(np.real(Pe) > 1, 1.0 - 1.0 / Pe, 0.0)
                    delta_k = hk / 2.0 / u_norm * Pe_f
                    f1[r] = f1[r] + (delta_k * _dot(fl, gC[k]))[..., None] * fl
                    if rterm is not None:
                        f1[r] = f1[r] - (rterm * delta_k)[..., None] * fl
            if self.flow["porous"]:                                         # volumetric Butler-Volmer, :1719-1726
                av = pp["specific surface area"] * self.Sw
                for echem in self.echem_params:
                    if cp["name"] in echem["stoichiometry"]:
                        nu = echem["stoichiometry"][cp["name"]]
                        f0[r] = f0[r] - av * nu / echem["electrons"] / pp["F"] * echem["reaction"](self.reaction_u(u), x)
        if self.flow["electroneutrality full"]:                             # :2015-2031, tested with v[num_liquid-1]
            rc = self.row_constraint
            for k in range(self.num_c):
                z = self.conc_params[k].get("z")
                if z != 0:
                    f0[rc] = f0[rc] + z * u[k]
        if self.i_Us is not None and self.echem_params:                     # solid Poisson source, :2162-2163
            av = pp["specific surface area"] * self.Sw
            for echem in self.echem_params:
                f0[self.i_Us] = f0[self.i_Us] + av * echem["reaction"](self.reaction_u(u), x)
        if self.flow["electroneutrality full"] and self.i_Us is not None:   # solid conduction, :2171
            f1[self.i_Us] = f1[self.i_Us] + self.solid_KU() * gu[self.i_Us]
        if self.flow["electroneutrality"]:
            iU = self.i_Ul
            for k in self.idx_c + [self.i_el]:
                cp = self.conc_params[k]
                z = cp["z"]
                if z == 0.0:
                    continue
                nd = cp.get("C_ND", 1.0)
                zDF = z * self.Dk(k) * pp["F"] * nd
                if self.flow["diffusion"]:
                    f1[iU] = f1[iU] + zDF * gC[k]                           # :1924
                if reactions is not None and not _is_zero(reactions[k]):
                    term = nd * z * pp["F"] * reactions[k]
                    f0[iU] = f0[iU] - term                                  # :1946
                    self.cancel0[iU] = self.cancel0[iU] + np.abs(np.real(term)) + 1j * np.abs(np.imag(term))
                if self.flow["porous"]:                                     # :1963-1969
                    av = pp["specific surface area"] * self.Sw
                    for echem in self.echem_params:
                        if cp["name"] in echem["stoichiometry"]:
                            nu = echem["stoichiometry"][cp["name"]]
                            f0[iU] = f0[iU] - nd * av * z * nu / echem["electrons"] * echem["reaction"](self.reaction_u(u), x)
            K_U = self.conductivity(C)
            f1[iU] = f1[iU] + _bc(K_U)[..., None] * gU                      # :1978
            if self.i_Us is not None:                                       # solid Poisson, :2162-2186
                iS = self.i_Us
                f1[iS] = f1[iS] + self.solid_KU() * gu[iS]
                n_r = self.num_liquid + 1
                if reactions is not None and len(reactions) > n_r and not _is_zero(reactions[n_r]):
                    f0[iS] = f0[iS] - reactions[n_r]
        elif self.flow["poisson"]:
            iU = self.i_Ul
            eps = pp.get("relative permittivity") is not None and pp.get("vacuum permittivity") is not None
            for k in range(self.num_liquid):
                cp = self.conc_params[k]
                z = cp.get("z")
                if z is None or z == 0:
                    continue
                if eps:
                    f0[iU] = f0[iU] - pp["F"] * z * C[k]                    # :2141
            K_U = self.poisson_KU(C)
            f1[iU] = f1[iU] + _bc(K_U)[..., None] * gU                      # :2171
            n_r = self.num_liquid                                           # :2178-2186 ("for test only")
            if reactions is not None and len(reactions) > n_r and not _is_zero(reactions[n_r]):
                f0[iU] = f0[iU] - reactions[n_r]
        return f0, f1

    # -- interior facets (DG) -------------------------------------------
    def interior_facet(self, x, n, up, gup, um, gum, geop, geom):
        """Coefficients for (v+, grad v+, v-, grad v-); n is the '+' outward normal."""
        nf = self.num_fields
        dt = up[0].dtype
        zero = np.zeros(x.shape[:-1], dtype=dt)
        zv = np.zeros(x.shape, dtype=dt)
        f0p = [zero.copy() for _ in range(nf)]
        f0m = [zero.copy() for _ in range(nf)]
        f1p = [zv.copy() for _ in range(nf)]
        f1m = [zv.copy() for _ in range(nf)]
        if self.family != "DG":
            return f0p, f1p, f0m, f1m
        pp = self.physical_params
        Cp, gCp = self.concentrations(up, gup)
        Cm, gCm = self.concentrations(um, gum)
        D_IPp = self.penalty(geop)
        D_IPm = self.penalty(geom)
        if self.has_potential:
            Up, gUp = up[self.i_Ul], gup[self.i_Ul]
            Um, gUm = um[self.i_Ul], gum[self.i_Ul]
        else:
            Up = gUp = Um = gUm = None

        def sip(i, kp, km, valp, valm, gp, gm, pen):
            """-avg(k grad w).jump(v,n) - avg(k grad v).jump(w,n) + pen jump(w) jump(v)."""
            kp_ = _bc(kp)
            km_ = _bc(km)
            a = -0.5 * (kp_ * _dot(gp, n) + km_ * _dot(gm, n)) + pen * (valp - valm)
            f0p[i] = f0p[i] + a
            f0m[i] = f0m[i] - a
            jw = (valp - valm)
            f1p[i] = f1p[i] - 0.5 * (kp_ * jw)[..., None] * n
            f1m[i] = f1m[i] - 0.5 * (km_ * jw)[..., None] * n

        for i, k in enumerate(self.idx_c):
            cp = self.conc_params[k]
            D = self.Dk(k)
            if self.flow["diffusion"]:
                pen = 0.5 * (D_IPp * D + D_IPm * D)                         # :1518-1520
                sip(i, D, D, Cp[k], Cm[k], gCp[k], gCm[k], pen)
            flp = self.flow_vector(k, x, gUp)
            if flp is not None:
                flm = self.flow_vector(k, x, gUm)
                fnp = _dot(flp, n)
                fnm = -_dot(flm, n)
                flow_Np = 0.5 * (fnp + _cabs(fnp))                          # :1652
                flow_Nm = 0.5 * (fnm + _cabs(fnm))
                a = flow_Np * Cp[k] - flow_Nm * Cm[k]                       # :1653-1654
                f0p[i] = f0p[i] + a
                f0m[i] = f0m[i] - a
        if self.flow["electroneutrality"]:
            iU = self.i_Ul
            for k in self.idx_c + [self.i_el]:
                cp = self.conc_params[k]
                z = cp["z"]
                if z == 0.0:
                    continue
                zDF = z * self.Dk(k) * pp["F"] * cp.get("C_ND", 1.0)
                if self.flow["diffusion"]:
                    sip(iU, zDF, zDF, Cp[k], Cm[k], gCp[k], gCm[k], 0.0)    # :1927-1930
            D_IP2p, D_IP2m = D_IPp, D_IPm
            KUp = self.conductivity(Cp)
            KUm = self.conductivity(Cm)
            pen = 0.5 * (D_IP2p * KUp + D_IP2m * KUm)                       # :1983
            sip(iU, KUp, KUm, Up, Um, gUp, gUm, pen)
            if self.i_Us is not None:                                       # :2171-2176
                iS = self.i_Us
                Ks = self.solid_KU()
                sip(iS, Ks, Ks, up[iS], um[iS], gup[iS], gum[iS], 0.5 * (D_IPp * Ks + D_IPm * Ks))
        elif self.flow["poisson"]:
            iU = self.i_Ul
            KUp = self.poisson_KU(Cp)
            KUm = self.poisson_KU(Cm)
            pen = 0.5 * (D_IPp * KUp + D_IPm * KUm)                         # :2174-2176
            sip(iU, KUp, KUm, Up, Um, gUp, gUm, pen)
        return f0p, f1p, f0m, f1m

    # -- exterior facets --------------------------------------------------
    def exterior_facet(self, marker, x, n, u, gu, geo, aux):
        """Coefficients (f0, f1) on an exterior facet carrying `marker`.

        aux["U_app"]: the applied potential as the reference sees it, i.e. the
        Vu-interpolant when U_app is an expression (solver.py:373-380).
        """
        nf = self.num_fields
        dt = u[0].dtype
        zero = np.zeros(x.shape[:-1], dtype=dt)
        f0 = [zero.copy() for _ in range(nf)]
        f1 = [np.zeros(x.shape, dtype=dt) for _ in range(nf)]
        pp = self.physical_params
        bm = self.boundary_markers
        DG = self.family == "DG"

        def on(name):
            ids = bm.get(name)
            return ids is not None and marker in ids

        C, gC = self.concentrations(u, gu)
        D_IP = self.penalty(geo) if DG else None
        if self.has_potential:
            U, gU = u[self.i_Ul], gu[self.i_Ul]
        else:
            U = gU = None
        ru = self.reaction_u(u)

        def nitsche(i, kcoef, w, gw, w0, pen):
            """k (w0 - w) n.grad v - k grad w.n v + pen k (w - w0) v  (solver.py:1525-1529)."""
            k_ = _bc(kcoef)
            f1[i] = f1[i] + (k_ * (w0 - w))[..., None] * n
            f0[i] = f0[i] - k_ * _dot(gw, n)
            if pen is not None:
                f0[i] = f0[i] + pen * k_ * (w - w0)

        applied_ids = None
        if self.flow["porous"]:
            if bm.get("liquid applied") is not None:
                applied_ids = bm["liquid applied"]                         # :1672-1680
        elif bm.get("applied") is not None:
            applied_ids = bm["applied"]

        for i, k in enumerate(self.idx_c):
            cp = self.conc_params[k]
            D = self.Dk(k)
            r = self.row[i]
            C_0 = _evalf(cp.get("bulk"), x) if cp.get("bulk") is not None else None
            if self.flow["diffusion"]:
                if on("bulk dirichlet") and DG:
                    nitsche(r, D, C[k], gC[k], C_0, D_IP)                   # :1523-1529
                if on("gas") and cp.get("gas") is not None and DG:
                    nitsche(r, D, C[k], gC[k], _evalf(cp["gas"], x), D_IP)  # :1533-1538
            fl = self.flow_vector(k, x, gU)
            if fl is not None:
                if applied_ids is not None and marker in applied_ids:       # :1681-1685
                    fn = _dot(fl, n)
                    f0[r] = f0[r] + _neg(fn) * fn * C_0 + _pos(fn) * fn * C[k]
                if bm.get("inlet") is not None:                             # :1687-1693
                    ids = set(bm["inlet"]) - set(applied_ids or ())
                    if marker in ids and not _is_zero(cp.get("bulk")):
                        vn = _dot(self.vel(x), n)
                        f0[r] = f0[r] + _neg(vn) * vn * C_0
                if bm.get("outlet") is not None:                            # :1695-1700
                    ids = set(bm["outlet"]) - set(applied_ids or ())
                    if marker in ids:
                        vn = _dot(self.vel(x), n)
                        f0[r] = f0[r] + _pos(vn) * vn * C[k]
            if on("bulk") and cp.get("mass transfer coefficient") is not None:
                f0[r] = f0[r] - cp["mass transfer coefficient"] * (C_0 - C[k])  # :1703-1706
            for echem in ([] if self.flow["porous"] else self.echem_params):   # :1710-1717 (surface reactions)
                if marker in bm.get(echem["boundary"], ()) and cp["name"] in echem["stoichiometry"]:
                    nu = echem["stoichiometry"][cp["name"]]
                    f0[r] = f0[r] - nu / echem["electrons"] / pp["F"] * echem["reaction"](ru, x)
            if on("neumann"):                                               # :1729-1731
                f0[r] = f0[r] - self.neumann(C[k], cp, ru, x)

        if self.flow["electroneutrality"]:
            iU = self.i_Ul
            for k in self.idx_c + [self.i_el]:
                cp = self.conc_params[k]
                z = cp["z"]
                if z == 0.0:
                    continue
                nd = cp.get("C_ND", 1.0)
                C_0 = _evalf(cp.get("bulk"), x) if cp.get("bulk") is not None else None
                zDF = z * self.Dk(k) * pp["F"] * nd
                if self.flow["diffusion"] and DG:
                    if on("bulk dirichlet"):
                        nitsche(iU, zDF, C[k], gC[k], C_0, None)           # :1933-1937
                    if on("gas") and cp.get("gas") is not None:
                        nitsche(iU, zDF, C[k], gC[k], _evalf(cp["gas"], x), None)
                if on("bulk") and cp.get("mass transfer coefficient") is not None:
                    f0[iU] = f0[iU] - nd * z * pp["F"] * cp["mass transfer coefficient"] * (C_0 - C[k])
                for echem in ([] if self.flow["porous"] else self.echem_params):   # :1955-1962
                    if marker in bm.get(echem["boundary"], ()) and cp["name"] in echem["stoichiometry"]:
                        nu = echem["stoichiometry"][cp["name"]]
                        f0[iU] = f0[iU] - nd * z * nu / echem["electrons"] * echem["reaction"](ru, x)
                if on("neumann"):
                    f0[iU] = f0[iU] - nd * z * pp["F"] * self.neumann(C[k], cp, ru, x)
            K_U = self.conductivity(C)
            diric = []
            if on("bulk"):
                diric.append(0.0)                                           # :1992-1993
            if on("applied") and not self.flow["porous"]:
                diric.append(aux["U_app"])                                  # :1994-1995
            if on("liquid applied"):
                diric.append(aux["U_app"])
            for U_0 in diric:
                if DG:
                    nitsche(iU, K_U, U, gU, U_0, D_IP)                      # :2002-2006
            if self.i_Us is not None:                                       # solid: applied > inlet (:2188-2212)
                iS = self.i_Us
                tgt = None
                if bm.get("applied") is not None:
                    tgt = "applied"
                elif bm.get("inlet") is not None:
                    tgt = "inlet"
                if tgt is not None and on(tgt) and DG:
                    nitsche(iS, self.solid_KU(), u[iS], gu[iS], aux["U_app"], D_IP)
                if on("applied solid current"):
                    f0[iS] = f0[iS] - pp["applied current density"]
        elif self.flow["poisson"]:
            iU = self.i_Ul
            K_U = self.poisson_KU(C)
            # solver.py:2188-2212: exactly one Dirichlet set, applied wins over bulk
            if bm.get("applied") is not None:
                if on("applied") and DG:
                    nitsche(iU, K_U, U, gU, aux["U_app"], D_IP)
            elif bm.get("bulk") is not None:
                if on("bulk") and DG:
                    nitsche(iU, K_U, U, gU, 0.0, D_IP)
            if on("applied liquid current"):
                f0[iU] = f0[iU] - pp["applied current density"]            # :2222-2223
            if on("robin"):
                f0[iU] = f0[iU] - pp["gap capacitance"] * (aux["U_app"] - U)  # :2225-2228
            if on("poisson neumann"):
                f0[iU] = f0[iU] - pp["surface charge density"]             # :2230-2232
        return f0, f1

    # -- strong Dirichlet sets for CG (solver.py:1531, 2008-2011, 2214-2217) --
    def dirichlet_sets(self):
        """List of (field, marker ids, value spec) for the DirichletBC objects.

        value spec: ("bulk", k) species k's bulk; ("U_app",); ("zero",).
        """
        bm = self.boundary_markers
        out = []
        for i, k in enumerate(self.idx_c):
            cp = self.conc_params[k]
            if self.flow["diffusion"]:
                if bm.get("bulk dirichlet") is not None:
                    out.append((i, bm["bulk dirichlet"], ("bulk", k)))
                if bm.get("gas") is not None and cp.get("gas") is not None:
                    out.append((i, bm["gas"], ("gas", k)))
            if any(self.flow["diffusion finite size_" + t] for t in ("BK", "CS", "SP", "SMPNP", "BMCSL")):
                if bm.get("bulk dirichlet") is not None:                       # solver.py:1563, 1575, 1587, 1607, 1635
                    out.append((i, bm["bulk dirichlet"], ("bulk", k)))
        if self.flow["electroneutrality"]:
            if bm.get("bulk") is not None:
                out.append((self.i_Ul, bm["bulk"], ("zero",)))
            if bm.get("applied") is not None and not self.flow["porous"]:
                out.append((self.i_Ul, bm["applied"], ("U_app",)))
            if bm.get("liquid applied") is not None:
                out.append((self.i_Ul, bm["liquid applied"], ("U_app",)))
            if self.i_Us is not None:
                if bm.get("applied") is not None:
                    out.append((self.i_Us, bm["applied"], ("U_app",)))
                elif bm.get("inlet") is not None:
                    out.append((self.i_Us, bm["inlet"], ("U_app",)))
        elif self.flow["poisson"]:
            if bm.get("applied") is not None:
                out.append((self.i_Ul, bm["applied"], ("U_app",)))
            elif bm.get("bulk") is not None:
                out.append((self.i_Ul, bm["bulk"], ("zero",)))
        elif self.flow["electroneutrality full"]:                           # :2033-2044
            is_applied = (bm.get("applied") is not None and not self.flow["porous"]) or \
                bm.get("liquid applied") is not None
            if is_applied:
                out.append((self.i_Ul, bm["applied"], ("U_app",)))
            elif bm.get("bulk") is not None:
                out.append((self.i_Ul, bm["bulk"], ("zero",)))
            if self.i_Us is not None:                                       # :2188-2212, solid
                if bm.get("applied") is not None:
                    out.append((self.i_Us, bm["applied"], ("U_app",)))
                elif bm.get("inlet") is not None:
                    out.append((self.i_Us, bm["inlet"], ("U_app",)))
        return out


def _is_zero(a):
    return a is None or (np.isscalar(a) and a == 0)


def _bc(a):
    """Make a scalar-or-array broadcastable against (e, q) arrays."""
    return np.asarray(a)
