"""Function spaces and discrete fields (host-side bookkeeping, device storage).

Stands in for the few Firedrake objects the reference touches directly:
`FunctionSpace` / `MixedFunctionSpace` (solver.py:1424-1459, 260-272),
`Function` with `.assign`, `.interpolate`, `.sub(i)`, `.subfunctions`
(solver.py:306, 571-582, 663), `split`, `TestFunctions`.

Layout: a mixed vector is field-blocked; inside a field,
DG dofs are cell-contiguous: index = field * N + cell * n + local_node.
Storage is one float64 torch tensor, on the GPU when there is one.
"""
import numpy as np
import sympy as sp
import torch

from . import symbolic as sym
from .tables import ElementTables


def default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class FunctionSpace:
    def __init__(self, mesh, family, degree=None, vfamily=None, vdegree=None, variant="spectral", name=None):
        fam = {"DQ": "DG", "DG": "DG", "Discontinuous Lagrange": "DG",
               "CG": "CG", "Q": "CG", "Lagrange": "CG", "P": "CG"}.get(family)
        if fam is None:
            raise NotImplementedError("element family %r" % (family,))
        if vdegree is not None and vdegree != degree:
            raise NotImplementedError("anisotropic degree on extruded meshes")
        self.mesh = mesh
        self.family = fam
        self.degree = degree
        self.tables = ElementTables(degree, fam, variant)
        self.n = (degree + 1) ** mesh.dim
        if fam == "DG":
            self.cell_nodes = None                       # implicit: cell * n + a
            self.num_nodes = mesh.num_cells * self.n
        else:
            self._build_cg_map()
        self._node_coords = None

    def dim(self):
        return self.num_nodes

    def num_sub_spaces(self):
        return 1

    def __mul__(self, other):
        return MixedFunctionSpace([self, other])

    # -- node coordinates (interpolation points) ----------------------------------
    def ref_nodes(self):
        d = self.mesh.dim
        g = np.meshgrid(*([self.tables.nodes] * d), indexing="ij")
        return np.stack([a.reshape(-1) for a in g], axis=1)            # (n, d)

    def cell_node_coords(self):
        """(nc, n, d) physical coordinates of every cell's nodes (multilinear map)."""
        d = self.mesh.dim
        xi = self.ref_nodes()
        N = np.ones((xi.shape[0], 2 ** d))
        for v in range(2 ** d):
            for m in range(d):
                bit = (v >> (d - 1 - m)) & 1
                N[:, v] *= xi[:, m] if bit else (1.0 - xi[:, m])
        return np.einsum("av,cvk->cak", N, self.mesh.cell_coords())

    def node_coords(self):
        if self._node_coords is None:
            xc = self.cell_node_coords()
            if self.family == "DG":
                self._node_coords = xc.reshape(-1, self.mesh.dim)
            else:
                out = np.zeros((self.num_nodes, self.mesh.dim))
                out[self.cell_nodes.reshape(-1)] = xc.reshape(-1, self.mesh.dim)
                self._node_coords = out
        return self._node_coords

    def _build_cg_map(self):
        xc = self.cell_node_coords().reshape(-1, self.mesh.dim)
        span = np.abs(xc).max() + 1.0
        key = np.round(xc / span * 2.0 ** 40).astype(np.int64)
        _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
        self.cell_nodes = inv.reshape(self.mesh.num_cells, self.n).astype(np.int32)
        self.num_nodes = int(len(first))


class MixedFunctionSpace:
    def __init__(self, spaces):
        flat = []
        for s in spaces:
            flat += s.spaces if isinstance(s, MixedFunctionSpace) else [s]
        self.spaces = flat
        self.mesh = flat[0].mesh
        N = flat[0].num_nodes
        if any(s.num_nodes != N or s.family != flat[0].family or s.degree != flat[0].degree for s in flat):
            raise NotImplementedError("mixed spaces must share one scalar element")
        self.scalar = flat[0]

    def num_sub_spaces(self):
        return len(self.spaces)

    def dim(self):
        return sum(s.num_nodes for s in self.spaces)

    def sub(self, i):
        return _SubSpace(self, i)

    def __mul__(self, other):
        return MixedFunctionSpace([self, other])


class _SubSpace:
    def __init__(self, parent, index):
        self.parent, self.index = parent, index


class Function:
    """Discrete field; `kind`/`index` give its symbol when it appears in a form."""

    def __init__(self, space, name=None, data=None, offset=0, _symbol=None):
        self.space = space
        self.name = name
        if isinstance(space, MixedFunctionSpace):
            self._n = space.dim()
        else:
            self._n = space.num_nodes
        if data is None:
            data = torch.zeros(self._n, dtype=torch.float64, device=default_device())
            offset = 0
        self._data = data
        self._off = offset
        self._symbol = _symbol          # ('u'|'a', index) once registered with a solver
        self._expr = None               # symbolic source of the last interpolate()

    def function_space(self):
        return self.space

    @property
    def dat(self):
        """`f.dat.data` / `f.dat.data_ro` (PyOP2 Dat of the reference's Functions): host copy of the nodal values."""
        arr = self.numpy().copy()

        class _Dat:
            data = arr
            data_ro = arr
        return _Dat

    # -- storage ------------------------------------------------------------------
    @property
    def tensor(self):
        return self._data[self._off:self._off + self._n]

    def numpy(self):
        return self.tensor.detach().cpu().numpy()

    def sub(self, i):
        if not isinstance(self.space, MixedFunctionSpace):
            if i != 0:
                raise IndexError(i)
            return self
        off = self._off + sum(s.num_nodes for s in self.space.spaces[:i])
        return Function(self.space.spaces[i], data=self._data, offset=off)

    @property
    def subfunctions(self):
        if not isinstance(self.space, MixedFunctionSpace):
            return (self,)
        return tuple(self.sub(i) for i in range(len(self.space.spaces)))

    def split(self):
        return self.subfunctions

    # -- assignment ---------------------------------------------------------------
    def assign(self, value):
        if isinstance(value, Function):
            self.tensor.copy_(value.tensor)
            self._expr = value._expr
        elif isinstance(value, _Scaled):
            self.tensor.copy_(value.fn.tensor * value.factor)
        elif isinstance(value, sym.Expr):
            self.tensor.fill_(float(_eval_constant(value)))
        else:
            self.tensor.fill_(float(value))
        return self

    def interpolate(self, expr):
        if isinstance(self.space, MixedFunctionSpace):
            raise NotImplementedError("interpolate into a mixed space")
        x = self.space.node_coords()
        vals = evaluate_expression(expr, x)
        self.tensor.copy_(torch.from_numpy(np.ascontiguousarray(vals)).to(self._data.device))
        self._expr = expr
        return self

    def __truediv__(self, s):
        return _Scaled(self, 1.0 / float(s))

    def __mul__(self, s):
        if isinstance(s, (int, float)):
            return _Scaled(self, float(s))
        return self._as_expr() * s

    __rmul__ = __mul__

    # -- symbolic use ---------------------------------------------------------------
    def _as_expr(self):
        if self._symbol is None:
            raise ValueError("Function is not registered as a coefficient of a solver")
        kind, j = self._symbol
        return sym.FieldRef(kind, j, self.space.mesh.dim)


class VectorFunctionSpace:
    """`VectorFunctionSpace(mesh, family, degree)`: d copies of one scalar space (the velocity space of the
    reference's flow solvers, echemfem/flow_solver.py:45-60)."""

    def __init__(self, mesh, family, degree=None, dim=None, **kw):
        self.scalar = family if isinstance(family, FunctionSpace) else FunctionSpace(mesh, family, degree, **kw)
        self.mesh = mesh
        self.value_size = int(dim or mesh.dim)

    def dim(self):
        return self.value_size * self.scalar.num_nodes


class VectorFunction:
    """Discrete vector field: `value_size` scalar `Function`s of one space.  This is what `FlowSolver.vel`
    is in the reference (flow_solver.py:114) and what `EchemSolver.set_velocity` may assign to `self.vel`
    (examples/simple_flow_battery.py:98-114): the solver registers the components as data coefficients."""

    def __init__(self, space, name=None):
        self.space = space
        self.name = name
        self.components = [Function(space.scalar, name="%s_%d" % (name or "v", k)) for k in range(space.value_size)]

    def function_space(self):
        return self.space

    def sub(self, k):
        return self.components[k]

    @property
    def subfunctions(self):
        return tuple(self.components)

    def interpolate(self, expr):
        comps = list(expr) if not isinstance(expr, VectorFunction) else expr.components
        if len(comps) != len(self.components):
            raise ValueError("expected %d components" % len(self.components))
        for f, e in zip(self.components, comps):
            if isinstance(e, Function):
                f.assign(e)
            else:
                f.interpolate(e)
        return self

    def assign(self, other):
        return self.interpolate(other)

    def rename(self, name):
        self.name = name

    def numpy(self):
        return np.stack([f.numpy() for f in self.components], axis=1)

    def __len__(self):
        return len(self.components)

    def __iter__(self):
        return iter(self.components)

    def __getitem__(self, k):
        return self.components[k]


class _Scaled:
    def __init__(self, fn, factor):
        self.fn, self.factor = fn, factor


def _eval_constant(e):
    v = sym.as_expr(e).v
    sub = {s: sym.Constant.lookup(s) for s in v.free_symbols if sym.Constant.is_constant_symbol(s)}
    return v.xreplace(sub)


def evaluate_expression(expr, x):
    """Evaluate a scalar symbolic expression of the coordinates at points x (npts, d)."""
    if isinstance(expr, (int, float)):
        return np.full(x.shape[0], float(expr))
    e = sym.as_expr(expr).v
    syms = sorted(e.free_symbols, key=str)
    args = []
    for s in syms:
        n = str(s)
        if n[0] == "x" and n[1:].isdigit():
            args.append(x[:, int(n[1:])])
        elif sym.Constant.is_constant_symbol(s):
            args.append(np.full(x.shape[0], sym.Constant.lookup(s)))
        else:
            raise ValueError("cannot interpolate an expression depending on %s" % s)
    f = sp.lambdify(syms, e, [{"rpow": np.power, "PosPart": lambda v: np.maximum(v, 0.0)}, "numpy"])
    out = f(*args) if syms else float(e)
    return np.broadcast_to(np.asarray(out, dtype=float), (x.shape[0],)).copy()


def split(u):
    return u._split_symbols()


def errornorm(exact, fn, norm_type="L2", degree_rise=3):
    """L2 norm of (exact - fn), as `firedrake.errornorm` in the reference's tests."""
    V = fn.space
    mesh = V.mesh
    d = mesh.dim
    from .tables import gauss_points, basis_table
    gp, gw = gauss_points(V.degree + degree_rise)
    L, _ = basis_table(V.tables.nodes, gp)
    g = np.meshgrid(*([np.arange(len(gp))] * d), indexing="ij")
    qi = np.stack([a.reshape(-1) for a in g], axis=1)                 # (nq, d)
    m = V.degree + 1
    bi = np.stack([a.reshape(-1) for a in np.meshgrid(*([np.arange(m)] * d), indexing="ij")], axis=1)
    phi = np.ones((qi.shape[0], V.n))
    w = np.ones(qi.shape[0])
    for k in range(d):
        phi *= L[qi[:, k]][:, bi[:, k]]
        w *= gw[qi[:, k]]
    xi = gp[qi]                                                       # (nq, d)
    X = mesh.cell_coords()
    N = np.ones((xi.shape[0], 2 ** d))
    dN = np.ones((xi.shape[0], 2 ** d, d))
    for v in range(2 ** d):
        for k in range(d):
            bit = (v >> (d - 1 - k)) & 1
            val = xi[:, k] if bit else 1.0 - xi[:, k]
            N[:, v] *= val
            for l in range(d):
                dN[:, v, l] *= (1.0 if bit else -1.0) if l == k else val
    x = np.einsum("qv,cvk->cqk", N, X)
    J = np.einsum("qvl,cvk->cqkl", dN, X)
    detJ = np.abs(np.linalg.det(J))
    vals = fn.numpy()
    if V.family == "DG":
        loc = vals.reshape(mesh.num_cells, V.n)
    else:
        loc = vals[V.cell_nodes]
    uh = loc @ phi.T
    ex = evaluate_expression(exact, x.reshape(-1, d)).reshape(uh.shape)
    return float(np.sqrt(np.einsum("q,cq,cq->", w, detJ, (uh - ex) ** 2)))
