"""UFL -> `echemfem_b200.symbolic` lowering: the piece a Firedrake-hosted binding needs.

When the REAL reference runs (Firedrake present), `EchemSolver.Form` is a `ufl.Form` over `Function`s of a mixed
space (echemfem/solver.py:306-308, 396-399, 714-718).  `lower_form(Form, u, v, ...)` walks it and returns the
`symbolic.Form` the rest of this package consumes (`compiler.CompiledForm` -> physics module -> C ABI), so that the
hot path can be bound WITHOUT re-deriving the weak form: same forms, same user closures, our kernels.

UFL is not installable in the build container, so this module is written against UFL's public API
(`ufl.algorithms.apply_algebra_lowering`, `ufl.corealg.multifunction.MultiFunction`, `ufl.algorithms.map_integrands`)
and is exercised by `tests/test_ufl_lowering.py` only where `import ufl` succeeds.  It covers the node types the
reference's forms use (~35 types): arithmetic, `Power`, `Abs`, `Sqrt`/`Exp`/`Ln`/trigonometric and
hyperbolic functions, `MinValue`/`MaxValue`, `Conditional` with `LT/GT/LE/GE/And/Or/Not`, `Coefficient` (unknowns and
data fields), `Argument` (test functions), `Constant`, `SpatialCoordinate`, `FacetNormal`, `CellVolume`, `FacetArea`,
`CellDiameter`/`Circumradius`, `Grad`, `Indexed`/`ListTensor`/`ComponentTensor`/`IndexSum` (after algebra lowering),
`PositiveRestricted`/`NegativeRestricted`.  Anything else raises `NotImplementedError` naming the node type.
"""
import numpy as np

from . import symbolic as S


def _require_ufl():
    try:
        import ufl  # noqa: F401
    except Exception as e:                      # pragma: no cover - depends on the host
        raise RuntimeError("ufl is not importable on this host: %r" % (e,))
    return __import__("ufl")


class FieldMap:
    """Which UFL terminals are which point quantities.

    unknowns:   list of scalar UFL expressions `split(u)[j]` (or `u.sub(j)`), one per field of the mixed space, in W order
    tests:      list of `split(v)[i]` / `TestFunctions(W)[i]`
    data:       {ufl Coefficient: Function of this package}  (U_app interpolant, u_old, discrete velocity components)
    constants:  {ufl Constant: symbolic.Constant}
    """

    def __init__(self, mixed_u, mixed_v, nfields, dim, data=None, constants=None):
        self.u, self.v, self.nf, self.dim = mixed_u, mixed_v, nfields, dim
        self.data = dict(data or {})
        self.constants = dict(constants or {})


def make_lowering(fields):
    """The MultiFunction that maps a (scalar or tensor valued) UFL expression to `symbolic.Expr`."""
    ufl = _require_ufl()
    from ufl.corealg.multifunction import MultiFunction
    from ufl.corealg.map_dag import map_expr_dag
    d = fields.dim

    def scalar(e):
        e = S.as_expr(e)
        if e.ufl_shape != ():
            raise NotImplementedError("expected a scalar, got shape %r" % (e.ufl_shape,))
        return e

    class Lower(MultiFunction):
        def __init__(self):
            MultiFunction.__init__(self)
            self._index_values = {}

        # ---- terminals -------------------------------------------------------------------------------------
        def zero(self, o):
            return S.Expr(np.zeros(o.ufl_shape, dtype=object)) if o.ufl_shape else S.Expr(0)

        def scalar_value(self, o):
            return S.Expr(float(o))

        int_value = scalar_value
        real_value = scalar_value
        float_value = scalar_value

        def identity(self, o):
            n = o.ufl_shape[0]
            return S.Expr(np.array(np.eye(n), dtype=object))

        def spatial_coordinate(self, o):
            return S.coordinate(d)

        def facet_normal(self, o):
            return S.facet_normal(d)

        def cell_volume(self, o):
            return S.cell_volume()

        def facet_area(self, o):
            return S.facet_area()

        def cell_diameter(self, o):
            return S.cell_diameter()

        circumradius = cell_diameter          # never used by the reference; Firedrake maps both to cell size measures

        def constant(self, o):                # firedrake.Constant / ufl.Constant
            c = fields.constants.get(o)
            if c is None:
                vals = getattr(o, "values", None)
                c = S.Constant(np.asarray(vals()).reshape(o.ufl_shape) if callable(vals) else 0.0)
                fields.constants[o] = c
            return c

        constant_value = constant

        def coefficient(self, o):
            if o is fields.u:                 # the mixed unknown itself: vector of its fields
                return S.as_vector([S.FieldRef("u", j, d) for j in range(fields.nf)])
            f = fields.data.get(o)
            if f is None:
                raise NotImplementedError("UFL coefficient %r is neither the unknown nor a registered data field" % (o,))
            return f._as_expr() if hasattr(f, "_as_expr") else S.as_expr(f)

        def argument(self, o):
            if o is fields.v or o.number() == 0:
                return S.as_vector([S.FieldRef("v", i, d) for i in range(fields.nf)])
            raise NotImplementedError("trial functions are not lowered (the Jacobian is derived from the residual form)")

        # ---- tensor algebra (after apply_algebra_lowering only index notation is left) ----------------------------
        def multi_index(self, o):
            return tuple(self._index_values[i] if not hasattr(i, "_value") else int(i) for i in o.indices())

        def indexed(self, o, a, ii):
            arr = a.v if isinstance(a.v, np.ndarray) else np.asarray(a.v, dtype=object)
            return S.Expr(arr[ii]) if ii else a

        def list_tensor(self, o, *ops):
            return S.Expr(np.array([S.as_expr(c).v for c in ops], dtype=object))

        def component_tensor(self, o):        # handled unevaluated: needs its free indices bound, see expr() below
            raise NotImplementedError("ComponentTensor reached bottom-up; use lower_expr")

        def index_sum(self, o):
            raise NotImplementedError("IndexSum reached bottom-up; use lower_expr")

        # ---- arithmetic -------------------------------------------------------------------------------------
        def sum(self, o, a, b):
            return a + b

        def product(self, o, a, b):
            return a * b

        def division(self, o, a, b):
            return a / b

        def power(self, o, a, b):
            return scalar(a) ** scalar(b)

        def abs(self, o, a):
            return abs(scalar(a))

        def sqrt(self, o, a):
            return S.sqrt(a)

        def exp(self, o, a):
            return S.exp(a)

        def ln(self, o, a):
            return S.ln(a)

        def sin(self, o, a):
            return S.sin(a)

        def cos(self, o, a):
            return S.cos(a)

        def tan(self, o, a):
            return S.tan(a)

        def sinh(self, o, a):
            return S.sinh(a)

        def cosh(self, o, a):
            return S.cosh(a)

        def tanh(self, o, a):
            return S.tanh(a)

        def sign(self, o, a):
            return S.sign(a)

        def min_value(self, o, a, b):
            return S.min_value(a, b)

        def max_value(self, o, a, b):
            return S.max_value(a, b)

        # ---- conditionals -----------------------------------------------------------------------------------
        def lt(self, o, a, b):
            return S.lt(a, b)

        def gt(self, o, a, b):
            return S.gt(a, b)

        def le(self, o, a, b):
            return S.le(a, b)

        def ge(self, o, a, b):
            return S.ge(a, b)

        def and_condition(self, o, a, b):
            return S.And(a, b)

        def or_condition(self, o, a, b):
            return S.Or(a, b)

        def not_condition(self, o, a):
            return S.Not(a)

        def conditional(self, o, c, a, b):
            return S.conditional(c, a, b)

        # ---- derivatives and restrictions -------------------------------------------------------------------------
        def grad(self, o, a):
            return S.grad(a)

        def div(self, o, a):
            return S.div(a)

        def positive_restricted(self, o, a):
            return a("+")

        def negative_restricted(self, o, a):
            return a("-")

        def expr(self, o, *ops):
            raise NotImplementedError("UFL node type %s is not lowered" % type(o).__name__)

    lower = Lower()

    def lower_expr(e):
        """Top-down wrapper: binds free indices of IndexSum / ComponentTensor, everything else bottom-up."""
        from ufl.classes import IndexSum, ComponentTensor, Indexed
        if isinstance(e, IndexSum):
            summand, mi = e.ufl_operands
            (idx,) = mi.indices()
            total = None
            for k in range(e.dimension()):
                lower._index_values[idx] = k
                term = lower_expr(summand)
                total = term if total is None else total + term
            lower._index_values.pop(idx, None)
            return total
        if isinstance(e, ComponentTensor):
            body, mi = e.ufl_operands
            shape = e.ufl_shape
            out = np.empty(shape, dtype=object)
            for comp in np.ndindex(shape):
                for idx, k in zip(mi.indices(), comp):
                    lower._index_values[idx] = k
                out[comp] = S.as_expr(lower_expr(body)).v
            for idx in mi.indices():
                lower._index_values.pop(idx, None)
            return S.Expr(out)
        if isinstance(e, Indexed):
            a, mi = e.ufl_operands
            comp = tuple(int(i) if not type(i).__name__ == "Index" else lower._index_values[i] for i in mi.indices())
            base = S.as_expr(lower_expr(a))
            arr = base.v if isinstance(base.v, np.ndarray) else np.asarray(base.v, dtype=object)
            return S.Expr(arr[comp])
        if e._ufl_is_terminal_:
            return map_expr_dag(lower, e)
        # generic operator: lower the operands top-down, then apply the handler
        ops = [lower_expr(c) for c in e.ufl_operands]
        handler = getattr(lower, e._ufl_handler_name_, None)
        if handler is None:
            raise NotImplementedError("UFL node type %s is not lowered" % type(e).__name__)
        return handler(e, *ops)

    return lower_expr


def lower_form(form, fields):
    """`ufl.Form` -> `symbolic.Form` (same integrals: cell / exterior_facet(_vert) / interior_facet(_vert|_horiz),
    subdomain ids and quadrature degrees kept).  `form` is the residual form of the reference
    (`NonlinearVariationalProblem(Form, u, bcs)`, solver.py:714)."""
    _require_ufl()
    from ufl.algorithms import apply_algebra_lowering
    form = apply_algebra_lowering.apply_algebra_lowering(form)
    lower_expr = make_lowering(fields)
    kinds = {"cell": "cell", "exterior_facet": "exterior_facet", "exterior_facet_vert": "exterior_facet",
             "interior_facet": "interior_facet", "interior_facet_vert": "interior_facet",
             "interior_facet_horiz": "interior_facet"}
    out = S.Form()
    for it in form.integrals():
        kind = kinds.get(it.integral_type())
        if kind is None:        # exterior_facet_top / _bottom are never integrated by the reference (solver.py:222)
            raise NotImplementedError("integral type %s" % it.integral_type())
        sid = it.subdomain_id()
        sid = None if sid in ("otherwise", "everywhere") else (tuple(sid) if isinstance(sid, (tuple, list)) else (sid,))
        degree = (it.metadata() or {}).get("quadrature_degree")
        measure = S.Measure(kind, sid, degree)
        out = out + S.as_expr(lower_expr(it.integrand())) * measure
    return out
