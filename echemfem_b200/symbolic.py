"""Symbolic weak-form layer (sympy-backed).

The reference describes its PDE in UFL and lets Firedrake differentiate and
compile it (echemfem/solver.py:306-308, 396-399, 714-718).  User physics
arrives as Python callables that build such expressions (`echem["reaction"]`,
`physical_params["bulk reaction"]`, `set_velocity`, `"bulk"`).
This module provides the small expression language those callables need,
backed by sympy scalars so that linear-form extraction and the Gateaux
derivative are plain symbolic differentiation.  Only *pointwise integrands*
are generated from it; every loop is hand-written CUDA (csrc/).

Expressions are tensors (shape (), (d,) or (d, d)) of sympy scalars over a
fixed alphabet of point quantities:

  x0..          spatial coordinate          n0..     facet normal ('+' side)
  u{j}, gu{j}_{k}     field j and its gradient   (+ `_p` / `_m` restricted copies)
  a{j}, ga{j}_{k}     auxiliary (data) field j   (+ restricted copies)
  v{i}, gv{i}_{k}     test function of equation i (+ restricted copies)
  cv, cd, fa          CellVolume, CellDiameter (restrictable), FacetArea
  k{m}                runtime constants (`Constant`), assignable without recompiling
"""
import numbers

import numpy as np
import sympy as sp

_RESTRICTED = {}          # base symbol -> {'+': sym, '-': sym}


def _restrictable(name):
    s = sp.Symbol(name, real=True)
    if s not in _RESTRICTED:
        _RESTRICTED[s] = {"+": sp.Symbol(name + "_p", real=True),
                          "-": sp.Symbol(name + "_m", real=True)}
    return s


def sym(name):
    return sp.Symbol(name, real=True)


class rpow(sp.Function):
    """Real power base**expo for a non-integer exponent.

    sympy treats x**1.5 of a real symbol as possibly complex, which leaks re()/im() into derivatives of
    |.| (e.g. the upwind flux with Bruggeman coefficients eps**1.5).  UFL's `Power` is real-valued;
    this function keeps that semantics."""
    nargs = 2
    is_real = True
    is_extended_real = True

    @classmethod
    def eval(cls, base, expo):
        if base.is_Number and expo.is_Number:
            return sp.Float(float(base) ** float(expo))
        if expo.is_Integer:
            return base ** expo
        return None

    def fdiff(self, argindex=1):
        base, expo = self.args
        if argindex == 1:
            return expo * rpow(base, expo - 1)
        return self * sp.log(base)


def _power(a, b):
    a, b = sp.sympify(a), sp.sympify(b)
    if b.is_Integer or (a.is_Number and b.is_Number):
        return a ** b
    if b.is_Number and not a.is_Number:
        if a.is_nonnegative:
            return a ** b
        return rpow(a, b)
    return a ** b


class Expr:
    """Tensor of sympy scalars with UFL-like operators."""
    __array_priority__ = 100

    def __init__(self, value):
        if isinstance(value, Expr):
            value = value.v
        if isinstance(value, (list, tuple)):
            value = np.array([as_expr(c).v for c in value], dtype=object)
        if isinstance(value, np.ndarray):
            if value.shape == ():
                value = value.item()
            else:
                out = np.empty(value.shape, dtype=object)
                for idx in np.ndindex(value.shape):
                    out[idx] = sp.sympify(value[idx])
                value = out
        if not isinstance(value, np.ndarray):
            value = sp.sympify(value)
        self.v = value

    # -- shape ------------------------------------------------------------
    @property
    def ufl_shape(self):
        return self.v.shape if isinstance(self.v, np.ndarray) else ()

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return Expr(self.v[i])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    # -- arithmetic ---------------------------------------------------------
    def _bin(self, other, op):
        if isinstance(other, Form):
            return NotImplemented
        a, b = self.v, as_expr(other).v
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            return Expr(np.frompyfunc(op, 2, 1)(_obj(a), _obj(b)))
        return Expr(op(a, b))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __pow__(self, o): return self._bin(o, _power)
    def __rpow__(self, o): return self._bin(o, lambda a, b: _power(b, a))
    def __neg__(self): return self._map(lambda e: -e)
    def __pos__(self): return self
    def __abs__(self): return Expr(sp.Abs(self.v))

    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form([Integral(self, o)])
        return self._bin(o, lambda a, b: a * b)

    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)

    # comparisons build conditions (UFL: `dot(flow, n) < 0`)
    def __lt__(self, o): return Cond(sp.StrictLessThan(self.v, as_expr(o).v))
    def __gt__(self, o): return Cond(sp.StrictGreaterThan(self.v, as_expr(o).v))
    def __le__(self, o): return Cond(sp.LessThan(self.v, as_expr(o).v))
    def __ge__(self, o): return Cond(sp.GreaterThan(self.v, as_expr(o).v))

    # -- restriction: expr('+'), expr('-') -----------------------------------
    def __call__(self, side):
        if side not in ("+", "-"):
            raise ValueError("restriction must be '+' or '-'")
        sub = {s: r[side] for s, r in _RESTRICTED.items()}
        if side == "-":
            for k in range(3):
                sub[sym("n%d" % k)] = -sym("n%d" % k)
        return self._map(lambda e: e.xreplace(sub))

    def _map(self, f):
        if isinstance(self.v, np.ndarray):
            out = np.empty(self.v.shape, dtype=object)
            for idx in np.ndindex(self.v.shape):
                out[idx] = f(self.v[idx])
            return Expr(out)
        return Expr(f(self.v))

    def free_symbols(self):
        if isinstance(self.v, np.ndarray):
            s = set()
            for idx in np.ndindex(self.v.shape):
                s |= self.v[idx].free_symbols
            return s
        return self.v.free_symbols

    def __repr__(self):
        return "Expr(%s)" % (self.v,)

    def __bool__(self):
        raise TypeError("symbolic expression has no truth value")

    def __eq__(self, o):          # `bulk_reaction_term != 0`, `C_0 != 0.0` (solver.py:1499, 1691)
        if isinstance(o, numbers.Number) and not isinstance(self.v, np.ndarray):
            return bool(self.v == o)
        return NotImplemented

    def __ne__(self, o):
        if isinstance(o, numbers.Number) and not isinstance(self.v, np.ndarray):
            return not bool(self.v == o)
        return NotImplemented

    __hash__ = object.__hash__


def _obj(a):
    if isinstance(a, np.ndarray):
        return a
    z = np.empty((), dtype=object)
    z[()] = a
    return z


class Cond:
    def __init__(self, rel):
        self.rel = rel


def as_expr(a):
    if isinstance(a, Expr):
        return a
    if hasattr(a, "_as_expr"):
        return a._as_expr()
    return Expr(a)


# -- runtime constants ---------------------------------------------------------
class Constant(Expr):
    """Runtime scalar/vector constant: a kernel argument, not a literal."""
    _registry = []

    def __init__(self, value, domain=None):
        arr = np.asarray(value, dtype=float)
        self._slots = []
        if arr.shape == ():
            self.v = self._new(float(arr))
        else:
            out = np.empty(arr.shape, dtype=object)
            for idx in np.ndindex(arr.shape):
                out[idx] = self._new(float(arr[idx]))
            self.v = out

    def _new(self, val):
        m = len(Constant._registry)
        s = sym("k%d" % m)
        Constant._registry.append([s, val])
        self._slots.append(m)
        return s

    def assign(self, value):
        arr = np.asarray(value.values() if isinstance(value, Constant) else value, dtype=float).ravel()
        for m, val in zip(self._slots, arr if arr.size == len(self._slots) else np.repeat(arr, len(self._slots))):
            Constant._registry[m][1] = float(val)
        return self

    def values(self):
        vals = np.array([Constant._registry[m][1] for m in self._slots])
        return vals if len(vals) > 1 else vals[0]

    def __float__(self):
        return float(self.values())

    @staticmethod
    def lookup(symbol):
        m = int(str(symbol)[1:])
        return Constant._registry[m][1]

    @staticmethod
    def is_constant_symbol(s):
        n = str(s)
        return n[0] == "k" and n[1:].isdigit()


# -- point quantities ------------------------------------------------------------
def coordinate(d):
    return Expr(np.array([sym("x%d" % k) for k in range(d)], dtype=object))


def facet_normal(d):
    return Expr(np.array([sym("n%d" % k) for k in range(d)], dtype=object))


def cell_volume():
    return Expr(_restrictable("cv"))


def cell_diameter():
    return Expr(_restrictable("cd"))


def facet_area():
    return Expr(sym("fa"))


class FieldRef(Expr):
    """Value of discrete field j ('u' = unknown, 'a' = auxiliary data, 'v' = test)."""

    def __init__(self, kind, j, d):
        self.kind, self.j, self.d = kind, j, d
        self.v = _restrictable("%s%d" % (kind, j))
        self._grad = np.array([_restrictable("g%s%d_%d" % (kind, j, k)) for k in range(d)], dtype=object)


def field_symbols(kind, j, d):
    return (_restrictable("%s%d" % (kind, j)),
            [_restrictable("g%s%d_%d" % (kind, j, k)) for k in range(d)])


def parse_field_symbol(s):
    """-> (kind, j, comp or None, side or None) for u/a/v symbols, else None."""
    name = str(s)
    side = None
    if name.endswith("_p"):
        side, name = "+", name[:-2]
    elif name.endswith("_m"):
        side, name = "-", name[:-2]
    comp = None
    if name[0] == "g" and len(name) > 1 and name[1] in "uav":
        body = name[1:]
        if "_" not in body:
            return None
        head, c = body.split("_")
        if not (head[1:].isdigit() and c.isdigit()):
            return None
        return head[0], int(head[1:]), int(c), side
    if name[0] in "uav" and name[1:].isdigit():
        return name[0], int(name[1:]), comp, side
    return None


# -- differential operators ----------------------------------------------------
def _grad_scalar(e, d):
    """Chain rule over the point alphabet: d e/dx_k + sum de/dw * grad(w)_k."""
    out = [sp.Integer(0)] * d
    for s in e.free_symbols:
        name = str(s)
        if name[0] == "x" and name[1:].isdigit():
            k = int(name[1:])
            out[k] = out[k] + sp.diff(e, s)
            continue
        pf = parse_field_symbol(s)
        if pf is None:
            if name[0] == "n" and name[1:].isdigit():
                raise NotImplementedError("grad of an expression containing the facet normal")
            continue            # constants, cell-wise constants
        kind, j, comp, side = pf
        if comp is not None:
            raise NotImplementedError("second derivatives of discrete fields are not supported")
        base, g = field_symbols(kind, j, d)
        de = sp.diff(e, s)
        for k in range(d):
            gk = g[k] if side is None else _RESTRICTED[g[k]][side]
            out[k] = out[k] + de * gk
    return out


def _dim_of(e):
    d = 0
    for s in as_expr(e).free_symbols():
        name = str(s)
        if name[0] == "x" and name[1:].isdigit():
            d = max(d, int(name[1:]) + 1)
        pf = parse_field_symbol(s)
        if pf is not None and pf[2] is not None:
            d = max(d, pf[2] + 1)
    return d


_GEOM_DIM = [2]


def set_geometric_dimension(d):
    _GEOM_DIM[0] = d


def grad(e):
    e = as_expr(e)
    d = _GEOM_DIM[0]
    if e.ufl_shape == ():
        return Expr(np.array(_grad_scalar(e.v, d), dtype=object))
    if len(e.ufl_shape) == 1:
        return Expr(np.array([_grad_scalar(c, d) for c in e.v], dtype=object))
    raise NotImplementedError("grad of rank-2 tensors")


def div(e):
    e = as_expr(e)
    g = grad(e).v
    return Expr(sum(g[k, k] for k in range(g.shape[0])))


def inner(a, b):
    a, b = as_expr(a), as_expr(b)
    if a.ufl_shape != b.ufl_shape:
        raise ValueError("inner: shape mismatch %s vs %s" % (a.ufl_shape, b.ufl_shape))
    if a.ufl_shape == ():
        return Expr(a.v * b.v)
    return Expr(sum(x * y for x, y in zip(a.v.ravel(), b.v.ravel())))


def dot(a, b):
    a, b = as_expr(a), as_expr(b)
    if a.ufl_shape == () or b.ufl_shape == ():
        return a * b
    return Expr(np.dot(a.v, b.v))


def outer(a, b):
    a, b = as_expr(a), as_expr(b)
    return Expr(np.outer(a.v, b.v))


def as_vector(comps):
    return Expr(np.array([as_expr(c).v for c in comps], dtype=object))


def avg(e):
    e = as_expr(e)
    return 0.5 * (e("+") + e("-"))


def jump(e, n=None):
    e = as_expr(e)
    if n is None:
        return e("+") - e("-")
    n = as_expr(n)
    if e.ufl_shape == ():
        return e("+") * n("+") + e("-") * n("-")
    return dot(e("+"), n("+")) + dot(e("-"), n("-"))


# -- pointwise functions -----------------------------------------------------------
def _fn(f):
    def g(e):
        return as_expr(e)._map(f)
    return g


exp = _fn(sp.exp)
ln = _fn(sp.log)
sqrt = _fn(sp.sqrt)
sin = _fn(sp.sin)
cos = _fn(sp.cos)
tan = _fn(sp.tan)
tanh = _fn(sp.tanh)
sinh = _fn(sp.sinh)
cosh = _fn(sp.cosh)
sign = _fn(sp.sign)
pi = sp.pi


class PosPart(sp.Function):
    """(x + |x|) / 2 kept as ONE node.  sympy distributes the 1/2 and expands x, after which x and |x| are evaluated
    along different association orders and a negative x leaves eps * |x| instead of the exact 0 the reference's
    `0.5*(fn + abs(fn))` gives -- enough to spoil the 1e-12 entry tolerance of blocks whose own scale is 1e5 times
    smaller than the advective flux.  Printed as fmax(x, 0), which equals (x + |x|) / 2 bit for bit; differentiates
    as UFL does (d|x| = sign(x) dx)."""
    nargs = 1

    @classmethod
    def eval(cls, x):
        if x.is_Number:
            return sp.Max(x, 0)

    def fdiff(self, argindex=1):
        return (1 + sp.sign(self.args[0])) / 2


def positive_part(e):
    """(e + |e|) / 2 (upwind flux of echemfem/solver.py:1652-1654)."""
    return as_expr(e)._map(PosPart)


def max_value(a, b):
    return Expr(sp.Max(as_expr(a).v, as_expr(b).v))


def min_value(a, b):
    return Expr(sp.Min(as_expr(a).v, as_expr(b).v))


def conditional(c, a, b):
    a, b = as_expr(a), as_expr(b)
    return Expr(sp.Piecewise((a.v, c.rel), (b.v, True)))


def gt(a, b): return as_expr(a) > b
def lt(a, b): return as_expr(a) < b
def ge(a, b): return as_expr(a) >= b
def le(a, b): return as_expr(a) <= b
def And(a, b): return Cond(sp.And(a.rel, b.rel))
def Or(a, b): return Cond(sp.Or(a.rel, b.rel))
def Not(a): return Cond(sp.Not(a.rel))


# -- measures and forms ----------------------------------------------------------
class Measure:
    def __init__(self, kind, subdomain_id=None, degree=None, weight=None):
        self.kind = kind                      # 'cell' | 'exterior_facet' | 'interior_facet'
        self.subdomain_id = subdomain_id
        self.degree = degree
        self.weight = weight                  # CylindricalMeasure: radius factor

    def __call__(self, subdomain_id=None, domain=None, scheme=None, degree=None, **kw):
        if isinstance(subdomain_id, int):
            subdomain_id = (subdomain_id,)
        elif subdomain_id is not None:
            subdomain_id = tuple(subdomain_id)
        return Measure(self.kind, subdomain_id if subdomain_id is not None else self.subdomain_id,
                       degree if degree is not None else self.degree, self.weight)

    def __rmul__(self, e):
        return Form([Integral(as_expr(e), self)])


dx = Measure("cell")
ds = Measure("exterior_facet")
dS = Measure("interior_facet")


class Integral:
    def __init__(self, integrand, measure):
        integrand = as_expr(integrand)
        if integrand.ufl_shape != ():
            raise ValueError("integrand must be scalar")
        if measure.weight is not None:
            integrand = integrand * measure.weight
        self.integrand = integrand.v
        self.kind = measure.kind
        self.subdomain_id = measure.subdomain_id
        self.degree = measure.degree


class Form:
    def __init__(self, integrals=()):
        self.integrals = list(integrals)

    def __add__(self, o):
        if isinstance(o, Form):
            return Form(self.integrals + o.integrals)
        if isinstance(o, numbers.Number) and o == 0:
            return self
        return NotImplemented

    __radd__ = __add__

    def __neg__(self):
        out = []
        for it in self.integrals:
            n = Integral.__new__(Integral)
            n.__dict__.update(it.__dict__)
            n.integrand = -it.integrand
            out.append(n)
        return Form(out)

    def __sub__(self, o):
        return self + (-o)

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, w):
        w = as_expr(w).v
        out = []
        for it in self.integrals:
            n = Integral.__new__(Integral)
            n.__dict__.update(it.__dict__)
            n.integrand = it.integrand * w
            out.append(n)
        return Form(out)

    __rmul__ = __mul__

    def __eq__(self, o):          # `a0 == 0` (solver.py:632)
        return Equation(self, o)

    __hash__ = object.__hash__


class Equation:
    def __init__(self, lhs, rhs):
        self.lhs, self.rhs = lhs, rhs
