"""From a symbolic weak form to the pointwise integrands of the CUDA kernels.

The reference obtains its Jacobian with `derivative(F, u)` and its element
kernels from TSFC (echemfem/solver.py:714-718).  Here the
same two steps are done symbolically, but only *down to the quadrature point*:

  integrand = sum_i ( f0_i v_i + f1_i . grad v_i )           (linear in the test functions)
  g0_ij = d f0_i / d u_j        g1_ij = d f0_i / d grad u_j
  g2_ij = d f1_i / d u_j        g3_ij = d f1_i / d grad u_j

for cell, exterior-facet (per boundary class) and interior-facet integrals
(own side '+', neighbour side '-').  The result is a C++ header of
`__device__` functions plus compile-time sparsity masks; every loop, gather,
basis contraction and CSR store is hand-written in `csrc/ech_dg.cuh`.

`abs` differentiates to `sign` and `conditional` piecewise, as in UFL.
"""
import hashlib

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter

from . import symbolic as S
from .tables import ElementTables


class _Printer(C99CodePrinter):
    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 1 < abs(int(e)) <= 4:
            # always parenthesised: sympy's Mul printer places a positive power in a denominator as is
            # (x * y**-4 prints as x/<y**4>), and an open product there would bind as ((x/y)*y)*y*y
            s = "(%s)" % "*".join(["(%s)" % self._print(b)] * abs(int(e)))
            return s if e > 0 else "(1.0/%s)" % s
        if e == sp.Rational(1, 2):
            return "sqrt(%s)" % self._print(b)
        if e == -sp.Rational(1, 2):
            return "(1.0/sqrt(%s))" % self._print(b)
        if e == -1:
            return "(1.0/(%s))" % self._print(b)
        return "pow((double)(%s), (double)(%s))" % (self._print(b), self._print(e))

    def _print_rpow(self, expr):
        return "pow((double)(%s), (double)(%s))" % (self._print(expr.args[0]), self._print(expr.args[1]))

    def _print_sign(self, expr):
        a = self._print(expr.args[0])
        return "(((%s) > 0.0) ? 1.0 : (((%s) < 0.0) ? -1.0 : 0.0))" % (a, a)

    def _print_Abs(self, expr):
        return "fabs(%s)" % self._print(expr.args[0])

    def _print_PosPart(self, expr):
        # (x + |x|) / 2 == fmax(x, 0) bit for bit, with x evaluated once (symbolic.PosPart)
        return "fmax((double)(%s), 0.0)" % self._print(expr.args[0])

    def _print_Max(self, expr):
        args = [self._print(a) for a in expr.args]
        out = args[0]
        for a in args[1:]:
            out = "fmax((double)(%s), (double)(%s))" % (out, a)
        return out

    def _print_Min(self, expr):
        args = [self._print(a) for a in expr.args]
        out = args[0]
        for a in args[1:]:
            out = "fmin((double)(%s), (double)(%s))" % (out, a)
        return out

    def _print_Integer(self, expr):
        return "%d.0" % int(expr)

    def _print_Rational(self, expr):
        return "(%d.0/%d.0)" % (expr.p, expr.q)

    def _print_Float(self, expr):
        return repr(float(expr))

    def _print_Heaviside(self, expr):
        return "(((%s) > 0.0) ? 1.0 : 0.0)" % self._print(expr.args[0])

    def _print_DiracDelta(self, expr):
        return "0.0"


_PR = _Printer()


def _ccode(e):
    return _PR.doprint(e)


def _sym_side(name):
    if name.endswith("_m"):
        return name[:-2], "m"
    if name.endswith("_p"):
        return name[:-2], ""
    return name, ""


class PointFunction:
    """One generated device function: named outputs <- expressions."""

    def __init__(self, outputs):
        self.outputs = [(t, sp.sympify(e)) for t, e in outputs]

    def emit(self, const_index, indent="    "):
        lines = []
        exprs = [e for _, e in self.outputs]
        if not exprs:
            return ""
        repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("t_"), order="none")
        used = set()
        for _, e in repl:
            used |= e.free_symbols
        for e in red:
            used |= e.free_symbols
        tmp = {s for s, _ in repl}
        for s in sorted(used - tmp, key=str):
            lines.append("%sconst double %s = %s;" % (indent, s, _bind(str(s), const_index)))
        # sin and cos of the same argument (manufactured solutions, rotating fields) share one sincos() call
        by_arg = {}
        for s, e in repl:
            if isinstance(e, (sp.sin, sp.cos)):
                by_arg.setdefault(e.args[0], {})[type(e)] = s
        paired = {}
        for arg, d in by_arg.items():
            if sp.sin in d and sp.cos in d:
                first = d[sp.sin] if [x for x, _ in repl].index(d[sp.sin]) < [x for x, _ in repl].index(d[sp.cos]) \
                    else d[sp.cos]
                paired[d[sp.sin]] = paired[d[sp.cos]] = (first, d[sp.sin], d[sp.cos], arg)
        for s, e in repl:
            if s in paired:
                first, ss, sc, arg = paired[s]
                if s == first:
                    lines.append("%sdouble %s, %s;" % (indent, ss, sc))
                    lines.append("%ssincos(%s, &%s, &%s);" % (indent, _ccode(arg), ss, sc))
                continue
            lines.append("%sconst double %s = %s;" % (indent, s, _ccode(e)))
        for (t, _), e in zip(self.outputs, red):
            lines.append("%s%s = %s;" % (indent, t, _ccode(e)))
        return "\n".join(lines)


def _bind(name, const_index):
    base, side = _sym_side(name)
    if base[0] == "x" and base[1:].isdigit():
        return "p.x[%s]" % base[1:]
    if base[0] == "n" and base[1:].isdigit():
        return "p.n[%s]" % base[1:]
    if base in ("cv", "cd"):
        return "p.%s%s" % (base, side)
    if base == "fa":
        return "p.fa"
    if base[:2] == "kc" and base[2:].isdigit():
        return "p.K[%d]" % const_index[name]
    pf = S.parse_field_symbol(sp.Symbol(base))
    if pf is not None:
        kind, j, comp, _ = pf
        arr = {"u": "u", "a": "a"}[kind]
        if comp is None:
            return "p.%s%s[%d]" % (arr, side, j)
        return "p.g%s%s[%d][%d]" % (arr, side, j, comp)
    raise ValueError("unknown symbol %s in a weak form" % name)


class CompiledForm:
    """Pointwise data of one weak form: expressions, masks, generated header."""

    def __init__(self, form, nf, naux, dim, degree, family, variant="spectral"):
        self.nf, self.naux, self.dim, self.p, self.family = nf, naux, dim, degree, family
        self.tables = ElementTables(degree, family, variant)
        self._split_integrals(form)
        self._extract()
        self._constants()
        self.header = self._emit()
        self.key = hashlib.sha1(self.header.encode()).hexdigest()[:16]

    # -- 1. group integrals --------------------------------------------------------
    def _split_integrals(self, form):
        cell = sp.Integer(0)
        inter = sp.Integer(0)
        ext_all = sp.Integer(0)
        ext_by_id = {}
        for it in form.integrals:
            if it.kind == "cell":
                if it.subdomain_id is not None:
                    raise NotImplementedError("cell subdomains")
                cell = cell + it.integrand
            elif it.kind == "interior_facet":
                inter = inter + it.integrand
            else:
                if it.subdomain_id is None:
                    ext_all = ext_all + it.integrand
                else:
                    for m in it.subdomain_id:
                        ext_by_id[m] = ext_by_id.get(m, sp.Integer(0)) + it.integrand
        # boundary classes: identical integrand <=> same class; class 0 = unmarked facets
        self.ext_classes = [ext_all]
        self.marker_class = {}
        for m in sorted(ext_by_id):
            e = ext_by_id[m] + ext_all
            for c, ec in enumerate(self.ext_classes):
                if e == ec:
                    self.marker_class[m] = c
                    break
            else:
                self.ext_classes.append(e)
                self.marker_class[m] = len(self.ext_classes) - 1
        self.cell_integrand = cell
        self.int_integrand = inter

    # -- 2. linear-form extraction and Gateaux derivative ------------------------
    def _tests(self, side=""):
        d, nf = self.dim, self.nf
        sfx = {"": "", "+": "_p"}[side]
        v = [sp.Symbol("v%d%s" % (i, sfx), real=True) for i in range(nf)]
        gv = [[sp.Symbol("gv%d_%d%s" % (i, k, sfx), real=True) for k in range(d)] for i in range(nf)]
        return v, gv

    def _trials(self, sfx=""):
        d, nf = self.dim, self.nf
        u = [sp.Symbol("u%d%s" % (j, sfx), real=True) for j in range(nf)]
        gu = [[sp.Symbol("gu%d_%d%s" % (j, k, sfx), real=True) for k in range(d)] for j in range(nf)]
        return u, gu

    def _linear(self, integrand, side):
        v, gv = self._tests(side)
        f0 = [sp.diff(integrand, v[i]) for i in range(self.nf)]
        f1 = [[sp.diff(integrand, gv[i][k]) for k in range(self.dim)] for i in range(self.nf)]
        for e in f0 + [c for r in f1 for c in r]:
            bad = [s for s in e.free_symbols if str(s).startswith(("v", "gv")) and S.parse_field_symbol(s)]
            if bad:
                raise ValueError("weak form is not linear in the test functions (%s)" % bad)
        return f0, f1

    def _gateaux(self, f0, f1, sfx):
        u, gu = self._trials(sfx)
        nf, d = self.nf, self.dim
        g0 = [[sp.diff(f0[i], u[j]) for j in range(nf)] for i in range(nf)]
        g1 = [[[sp.diff(f0[i], gu[j][l]) for l in range(d)] for j in range(nf)] for i in range(nf)]
        g2 = [[[sp.diff(f1[i][k], u[j]) for k in range(d)] for j in range(nf)] for i in range(nf)]
        g3 = [[[[sp.diff(f1[i][k], gu[j][l]) for l in range(d)] for k in range(d)] for j in range(nf)]
              for i in range(nf)]
        return {"g0": g0, "g1": g1, "g2": g2, "g3": g3}

    def _extract(self):
        self.cell = {}
        self.cell["f0"], self.cell["f1"] = self._linear(self.cell_integrand, "")
        self.cell["J"] = self._gateaux(self.cell["f0"], self.cell["f1"], "")
        self.ext = []
        for e in self.ext_classes:
            f0, f1 = self._linear(e, "")
            self.ext.append({"f0": f0, "f1": f1, "J": self._gateaux(f0, f1, "")})
        self.int = {}
        self.int["f0"], self.int["f1"] = self._linear(self.int_integrand, "+")
        self.int["Jp"] = self._gateaux(self.int["f0"], self.int["f1"], "_p")
        self.int["Jm"] = self._gateaux(self.int["f0"], self.int["f1"], "_m")
        # field-block masks of the Jacobian: own-cell and neighbour blocks
        nf = self.nf
        own = np.eye(nf, dtype=bool)
        nbr = np.zeros((nf, nf), dtype=bool)
        for J in [self.cell["J"]] + [e["J"] for e in self.ext] + [self.int["Jp"]]:
            own |= _block_nonzero(J, nf)
        nbr |= _block_nonzero(self.int["Jm"], nf)
        own |= nbr
        self.own_mask, self.nbr_mask = own, nbr

    # -- 3. runtime constants -------------------------------------------------------
    def _all_exprs(self):
        out = []
        for grp in [self.cell] + self.ext + [self.int]:
            out += list(grp["f0"]) + [c for r in grp["f1"] for c in r]
            for key in ("J", "Jp", "Jm"):
                if key in grp:
                    J = grp[key]
                    out += _flatten(J["g0"]) + _flatten(J["g1"]) + _flatten(J["g2"]) + _flatten(J["g3"])
        return out

    def _constants(self):
        """Collect the runtime constants and rename them canonically (kc0, kc1, ...) so that the
        generated header -- and hence the module cache key -- does not depend on how many
        `Constant`s the process created before."""
        names = set()
        for e in self._all_exprs():
            for s in e.free_symbols:
                if S.Constant.is_constant_symbol(s):
                    names.add(str(s))
        self.const_source = sorted(names, key=lambda n: int(n[1:]))
        self.const_names = ["kc%d" % i for i in range(len(self.const_source))]
        self.const_index = {n: i for i, n in enumerate(self.const_names)}
        sub = {sp.Symbol(a, real=True): sp.Symbol(b, real=True) for a, b in zip(self.const_source, self.const_names)}

        def ren(x):
            if isinstance(x, list):
                return [ren(y) for y in x]
            return x.xreplace(sub)
        for grp in [self.cell] + self.ext + [self.int]:
            for key in list(grp.keys()):
                if key in ("J", "Jp", "Jm"):
                    grp[key] = {k: ren(v) for k, v in grp[key].items()}
                else:
                    grp[key] = ren(grp[key])

    def constant_values(self):
        return np.array([S.Constant.lookup(sp.Symbol(n)) for n in self.const_source] or [0.0])

    # -- 4. code generation -----------------------------------------------------------
    def _emit(self):
        nf, d, p = self.nf, self.dim, self.p
        T = self.tables
        L = []
        L.append("// generated by echemfem_b200.compiler -- pointwise integrands only; loops are in ech_dg.cuh")
        L.append("#pragma once")
        L.append("#define ECH_D %d" % d)
        L.append("#define ECH_P %d" % p)
        L.append("#define ECH_NF %d" % nf)
        L.append("#define ECH_NAUX %d" % self.naux)
        L.append("#define ECH_NA1 %d" % max(self.naux, 1))
        L.append("#define ECH_NCONST %d" % max(len(self.const_names), 1))
        L.append("#define ECH_NQ %d" % T.q)
        L.append("#define ECH_NCLS %d" % len(self.ext_classes))
        L.append("#define ECH_FAMILY_DG %d" % (1 if self.family == "DG" else 0))
        L.append("namespace echgen {")
        L.append("constexpr int D = ECH_D, P = ECH_P, NF = ECH_NF, NQ = ECH_NQ, N1 = ECH_P + 1;")
        L.append(_carray("TAB_L", T.L))
        L.append(_carray("TAB_DL", T.dL))
        L.append(_carray("TAB_W", T.gw))
        L.append(_carray("TAB_X", T.pts))
        # compile-time copies: reference bases of own-side points become immediates in ech_dg_coop2.cuh
        L.append(_carray("CT_L", T.L, "constexpr"))
        L.append(_carray("CT_DL", T.dL, "constexpr"))
        L.append("struct Pt { double x[D], n[D], u[NF], gu[NF][D], um[NF], gum[NF][D], "
                 "a[ECH_NA1], ga[ECH_NA1][D], am[ECH_NA1], gam[ECH_NA1][D], cv, cvm, cd, cdm, fa; "
                 "const double* __restrict__ K; };")
        L.append("struct Res { double f0[NF]; double f1[NF][D]; };")
        L.append("struct RowJ { double g0[NF]; double g1[NF][D]; double g2[NF][D]; double g3[NF][D][D]; };")
        # masks
        ext_union = _union_groups(self.ext, nf, d)
        self.masks = {
            "C": _masks(self.cell["f0"], self.cell["f1"], self.cell["J"], nf, d),
            "E": ext_union,
            "IP": _masks(self.int["f0"], self.int["f1"], self.int["Jp"], nf, d),
            "IM": _masks(self.int["f0"], self.int["f1"], self.int["Jm"], nf, d),
        }
        for tag, m in self.masks.items():
            L.append("struct M%s {" % tag)
            for key in ("F0", "F1"):
                L.append("  static constexpr bool %s[NF] = {%s};" % (key, ",".join(str(int(b)) for b in m[key])))
            for key in ("G0", "G1", "G2"):
                L.append("  static constexpr bool %s[NF][NF] = {%s};" % (
                    key, ",".join("{" + ",".join(str(int(b)) for b in row) + "}" for row in m[key])))
            L.append("  static constexpr int G3[NF][NF] = {%s};" % (
                ",".join("{" + ",".join(str(int(b)) for b in row) + "}" for row in m["G3"])))
            L.append("};")
        L.append("constexpr bool OWN_BLOCK[NF][NF] = {%s};" % ",".join(
            "{" + ",".join(str(int(b)) for b in row) + "}" for row in self.own_mask))
        L.append("constexpr bool NBR_BLOCK[NF][NF] = {%s};" % ",".join(
            "{" + ",".join(str(int(b)) for b in row) + "}" for row in self.nbr_mask))
        ci = self.const_index
        # residual point functions
        L.append("__device__ __forceinline__ void cell_res(const Pt& p, Res& o) {")
        L.append(PointFunction(_res_outputs(self.cell["f0"], self.cell["f1"], self.masks["C"], "o")).emit(ci))
        L.append("}")
        L.append("__device__ __forceinline__ void ifacet_res(const Pt& p, Res& o) {")
        L.append(PointFunction(_res_outputs(self.int["f0"], self.int["f1"], self.masks["IP"], "o")).emit(ci))
        L.append("}")
        L.append("__device__ __forceinline__ void efacet_res(int cls, const Pt& p, Res& o) {")
        L.append("  switch (cls) {")
        for c, grp in enumerate(self.ext):
            L.append("  case %d: {" % c)
            L.append(PointFunction(_res_outputs(grp["f0"], grp["f1"], ext_union, "o")).emit(ci, "      "))
            L.append("  } break;")
        L.append("  default: break;")
        L.append("  }")
        L.append("}")
        # Jacobian rows
        L.append("template <int I> __device__ __forceinline__ void cell_jac(const Pt& p, RowJ& o);")
        L.append("template <int I> __device__ __forceinline__ void ifacet_jac(const Pt& p, RowJ& op, RowJ& om);")
        L.append("template <int I> __device__ __forceinline__ void efacet_jac(int cls, const Pt& p, RowJ& o);")
        for i in range(nf):
            L.append("template <> __device__ __forceinline__ void cell_jac<%d>(const Pt& p, RowJ& o) {" % i)
            L.append(PointFunction(_jac_outputs(self.cell["J"], self.masks["C"], i, "o", nf, d)).emit(ci))
            L.append("}")
            L.append("template <> __device__ __forceinline__ void ifacet_jac<%d>(const Pt& p, RowJ& op, RowJ& om) {" % i)
            outs = _jac_outputs(self.int["Jp"], self.masks["IP"], i, "op", nf, d) + \
                _jac_outputs(self.int["Jm"], self.masks["IM"], i, "om", nf, d)
            L.append(PointFunction(outs).emit(ci))
            L.append("}")
            L.append("template <> __device__ __forceinline__ void efacet_jac<%d>(int cls, const Pt& p, RowJ& o) {" % i)
            L.append("  switch (cls) {")
            for c, grp in enumerate(self.ext):
                L.append("  case %d: {" % c)
                L.append(PointFunction(_jac_outputs(grp["J"], ext_union, i, "o", nf, d)).emit(ci, "      "))
                L.append("  } break;")
            L.append("  default: break;")
            L.append("  }")
            L.append("}")
        L.append("}  // namespace echgen")
        return "\n".join(l for l in L if l != "") + "\n"


# -- helpers ---------------------------------------------------------------------------
def _flatten(a):
    out = []
    for x in a:
        if isinstance(x, list):
            out += _flatten(x)
        else:
            out.append(x)
    return out


def _nz(e):
    return e != 0


def _block_nonzero(J, nf):
    m = np.zeros((nf, nf), dtype=bool)
    for i in range(nf):
        for j in range(nf):
            m[i, j] = (_nz(J["g0"][i][j]) or any(_nz(c) for c in J["g1"][i][j])
                       or any(_nz(c) for c in J["g2"][i][j]) or any(_nz(c) for c in _flatten(J["g3"][i][j])))
    return m


def _g3_kind(blk, d):
    """0: zero, 1: isotropic (s * identity), 2: general."""
    if not any(_nz(c) for c in _flatten(blk)):
        return 0
    iso = all((blk[k][l] == 0) if k != l else (sp.expand(blk[k][k] - blk[0][0]) == 0)
              for k in range(d) for l in range(d))
    return 1 if iso else 2


def _masks(f0, f1, J, nf, d):
    return {
        "F0": [_nz(f0[i]) for i in range(nf)],
        "F1": [any(_nz(c) for c in f1[i]) for i in range(nf)],
        "G0": [[_nz(J["g0"][i][j]) for j in range(nf)] for i in range(nf)],
        "G1": [[any(_nz(c) for c in J["g1"][i][j]) for j in range(nf)] for i in range(nf)],
        "G2": [[any(_nz(c) for c in J["g2"][i][j]) for j in range(nf)] for i in range(nf)],
        "G3": [[_g3_kind(J["g3"][i][j], d) for j in range(nf)] for i in range(nf)],
    }


def _union_groups(groups, nf, d):
    ms = [_masks(g["f0"], g["f1"], g["J"], nf, d) for g in groups]
    out = ms[0]
    for m in ms[1:]:
        for key in ("F0", "F1"):
            out[key] = [a or b for a, b in zip(out[key], m[key])]
        for key in ("G0", "G1", "G2"):
            out[key] = [[a or b for a, b in zip(ra, rb)] for ra, rb in zip(out[key], m[key])]
        g3 = []
        for ra, rb in zip(out["G3"], m["G3"]):
            row = []
            for a, b in zip(ra, rb):
                row.append(max(a, b) if (a == b or 0 in (a, b)) else 2)
            g3.append(row)
        out["G3"] = g3
    return out


def _res_outputs(f0, f1, mask, o):
    outs = []
    for i in range(len(f0)):
        if mask["F0"][i]:
            outs.append(("%s.f0[%d]" % (o, i), f0[i]))
        if mask["F1"][i]:
            for k in range(len(f1[i])):
                outs.append(("%s.f1[%d][%d]" % (o, i, k), f1[i][k]))
    return outs


def _jac_outputs(J, mask, i, o, nf, d):
    outs = []
    for j in range(nf):
        if mask["G0"][i][j]:
            outs.append(("%s.g0[%d]" % (o, j), J["g0"][i][j]))
        if mask["G1"][i][j]:
            for l in range(d):
                outs.append(("%s.g1[%d][%d]" % (o, j, l), J["g1"][i][j][l]))
        if mask["G2"][i][j]:
            for k in range(d):
                outs.append(("%s.g2[%d][%d]" % (o, j, k), J["g2"][i][j][k]))
        kind = mask["G3"][i][j]
        if kind == 1:
            outs.append(("%s.g3[%d][0][0]" % (o, j), J["g3"][i][j][0][0]))
        elif kind == 2:
            for k in range(d):
                for l in range(d):
                    outs.append(("%s.g3[%d][%d][%d]" % (o, j, k, l), J["g3"][i][j][k][l]))
    return outs


def _carray(name, a, qual="__constant__"):
    a = np.asarray(a, dtype=float)
    if a.ndim == 1:
        return "%s double %s[%d] = {%s};" % (qual, name, a.shape[0], ", ".join(repr(float(x)) for x in a))
    rows = ["{" + ", ".join(repr(float(x)) for x in r) + "}" for r in a]
    return "%s double %s[%d][%d] = {%s};" % (qual, name, a.shape[0], a.shape[1], ", ".join(rows))
