"""Host-side logic of the product: mesh connectivity, tables, symbolic layer, contract validation."""
import numpy as np
import pytest

from echemfem_b200 import symbolic as S
from echemfem_b200.fd import *  # noqa: F401,F403
from echemfem_b200 import EchemSolver
from echemfem_b200.mesh import TensorMesh
from echemfem_b200.tables import ElementTables, lobatto_nodes


def test_connectivity_matches_oracle_mesh():
    from oracle.mesh import tensor_mesh
    ax = [np.linspace(0, 1, 4), np.linspace(0, 2, 3), np.linspace(0, 1, 3)]
    m = TensorMesh(ax)
    o = tensor_mesh(ax)
    assert m.num_cells == o.num_cells
    assert np.array_equal(m.nbr_cell, o.nbr)
    mask = o.nbr >= 0
    assert np.array_equal(m.nbr_lf[mask], o.nbr_lf[mask])
    assert m.num_interior_facets == o.int_facets.shape[0]
    assert np.all(m.nbr_orient == 0)               # tensor meshes are consistently oriented
    for f, (c, lf) in enumerate(o.ext_facets):
        assert m.ext_marker[c, lf] == o.ext_marker[f]


def test_connectivity_slots_are_sorted_ranks():
    m = UnitSquareMesh(3, 3)
    nbr, code, slot, slot_self, ncol = m.connectivity({1: 1, 2: 1, 3: 2, 4: 2})
    for c in range(m.num_cells):
        cells = {int(slot_self[c]): c}
        for lf in range(4):
            if nbr[c, lf] >= 0:
                cells[int(slot[c, lf])] = int(nbr[c, lf])
        assert sorted(cells) == list(range(ncol[c]))
        assert [cells[k] for k in sorted(cells)] == sorted(cells.values())
    assert set(np.unique(nbr[nbr < 0])) == {-2, -3}        # boundary classes 1 and 2


def test_tables_against_oracle_fem():
    from oracle import fem
    for p in (1, 2, 3):
        for fam in ("DG", "CG"):
            t = ElementTables(p, fam)
            assert np.allclose(t.nodes, fem.nodes_1d(p, fam), atol=1e-15)
            L, dL = fem.lagrange_1d(fem.nodes_1d(p, fam), t.pts)
            assert np.allclose(t.L, L, atol=1e-13) and np.allclose(t.dL, dL, atol=1e-12)
            assert t.q == fem.num_quad_points(p)
    assert np.allclose(lobatto_nodes(3), [0, 0.5, 1])


def test_symbolic_restriction_and_grad():
    S.set_geometric_dimension(2)
    u = S.FieldRef("u", 0, 2)
    v = S.FieldRef("v", 0, 2)
    n = S.facet_normal(2)
    j = jump(v, n)
    assert str(j.v[0]) in ("n0*v0_p - n0*v0_m", "-n0*v0_m + n0*v0_p")
    g = grad(u * u)
    assert g.v[1] == 2 * S.sym("u0") * S.sym("gu0_1")
    x = S.coordinate(2)
    assert div(as_vector((x[0] ** 2, x[0] * x[1]))).v == 3 * S.sym("x0")
    k = Constant(2.0)
    assert float(k) == 2.0
    k.assign(5.0)
    assert float(k) == 5.0
    with pytest.raises(NotImplementedError):
        grad(grad(u)[0])


class _Bad(EchemSolver):
    def __init__(self, flow, family="DG"):
        mesh = UnitSquareMesh(2, 2)
        cp = [{"name": "A", "diffusion coefficient": 1.0, "z": 1.0, "bulk": 1.0},
              {"name": "B", "diffusion coefficient": 1.0, "z": -1.0, "bulk": 1.0}]
        pp = {"flow": flow, "F": 1.0, "R": 1.0, "T": 1.0}
        super().__init__(cp, pp, mesh, family=family)

    def set_boundary_markers(self):
        self.boundary_markers = {"bulk": 1}


@pytest.mark.parametrize("flow,msg", [
    (["diffusion", "electroneutrality"], "Electroneutrality constraint requires Migration"),
    (["diffusion", "migration"], "Migration requires Electroneutrality or Poisson"),
    (["diffusion", "migration", "poisson", "electroneutrality"], "Electroneutrality not compatible with Poisson"),
    (["diffusion", "finite size"], "finite size effects only implemented with CG"),
    (["diffusion", "darcy"], "Darcy requires a porous model"),
])
def test_invalid_flow_combinations_raise(flow, msg):
    """Same NotImplementedError messages as the reference (echemfem/solver.py:115-142)."""
    with pytest.raises(NotImplementedError, match=msg):
        _Bad(flow)


def test_contract_attributes():
    import product_cases
    s = product_cases.AdvectionDiffusionMigrationSolver(2, 1.0)
    assert s.num_c == 2 and s.num_liquid == 1 and s.num_mass == 1 and s.i_el == 1
    assert s.i_c == {"C1": 0} and s.i_Ul == 1 and s.idx_c == [0]
    assert s.boundary_markers["applied"] == (1, 2, 3, 4)
    assert s.solver_parameters["pc_type"] == "lu" and s.solver_parameters["snes_max_it"] == 50
    s.init_solver_parameters(pc_type="ilu")
    assert s.solver_parameters["ksp_type"] == "fgmres" and s.solver_parameters["ksp_gmres_restart"] == 1000
    s.init_solver_parameters(pc_type="gmg")                                  # solver.py:1190-1201
    assert s.solver_parameters["pc_type"] == "mg" and s.solver_parameters["ksp_rtol"] == 1e-2
    with pytest.raises(NotImplementedError):
        s.init_solver_parameters(pc_type="amg")
    assert s.W.dim() == 2 * 16 and len(s.u.subfunctions) == 2


def test_box_aggregates_cover_every_dof_and_field():
    from echemfem_b200.coarse import box_aggregates
    g = (np.arange(20) + 0.5) / 20.0
    X, Y = np.meshgrid(g, 2.0 * g, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    agg, nc = box_aggregates(coords, 3, 48)
    n = coords.shape[0]
    assert agg.shape == (3 * n,) and agg.dtype == np.int32
    nagg = nc // 3
    assert nc == 3 * nagg and 8 <= nagg <= 32
    # fields use disjoint coarse dofs with the same box structure
    assert np.array_equal(agg[n:2 * n], agg[:n] + nagg) and np.array_equal(agg[2 * n:], agg[:n] + 2 * nagg)
    assert set(agg[:n]) == set(range(nagg))
    # boxes are spatially compact: every aggregate spans less than half of the domain in each direction
    for a in range(nagg):
        pts = coords[agg[:n] == a]
        assert (pts.max(axis=0) - pts.min(axis=0) < np.array([0.5, 1.0])).all()
    # degenerate inputs
    agg1, nc1 = box_aggregates(np.zeros((5, 2)), 2, 100)
    assert nc1 == 2 and set(agg1) == {0, 1}


def test_boundary_layer_meshes_follow_reference_node_positions():
    """echemfem/utility_meshes.py:9-195: uniform core + uniform layers; markers 1..4 as on box meshes."""
    from echemfem_b200 import RectangleBoundaryLayerMesh, IntervalBoundaryLayerMesh
    # examples/gupta_advection_migration.py-style: layers at y = 0 and y = Ly
    m = RectangleBoundaryLayerMesh(10, 6, 2.0, 1.0, 4, 0.1, boundary=(3, 4))
    assert m.shape == (10, 14) and m.num_cells == 140
    ys = m.axes[1]
    assert np.allclose(ys[:5], [0.0, 0.025, 0.05, 0.075, 0.1]) and np.allclose(ys[-5:], [0.9, 0.925, 0.95, 0.975, 1.0])
    assert np.allclose(np.diff(ys[4:11]), 0.8 / 6) and np.allclose(np.diff(m.axes[0]), 0.2)
    assert set(np.unique(m.ext_marker)) == {0, 1, 2, 3, 4}
    # different layer in x and y
    m2 = RectangleBoundaryLayerMesh(4, 4, 1.0, 1.0, 2, 0.2, Ly_bdlayer=0.1, ny_bdlayer=5, boundary=(1, 4))
    assert m2.shape == (6, 9)
    assert np.allclose(m2.axes[0][:3], [0.0, 0.1, 0.2]) and np.allclose(m2.axes[1][-6:], 0.9 + 0.02 * np.arange(6))
    # interval: default layer on the right only (boundary=(2,))
    mi = IntervalBoundaryLayerMesh(10, 1.0, 5, 0.1)
    x = mi.axes[0]
    assert mi.shape == (15,) and np.allclose(np.diff(x[:11]), 0.09) and np.allclose(np.diff(x[10:]), 0.02)
    assert mi.num_exterior_facets == 2 and set(np.unique(mi.ext_marker)) == {0, 1, 2}
    with pytest.raises(ValueError):
        IntervalBoundaryLayerMesh(10, 0.1, 5, 0.2, boundary=(1, 2))


def test_code_printer_keeps_integer_powers_together():
    """x / y**4 must not print as x/(y)*(y)*(y)*(y) (found by the BMCSL parity test)."""
    import sympy as sp
    from echemfem_b200.compiler import _Printer
    x, y = sp.symbols("x y")
    for expr in (x / y ** 4, x * y ** 3 / (1 - y) ** 2, y ** -3, (x + 1) ** 2 / y ** 2):
        code = _Printer().doprint(expr)
        val = eval(code.replace("pow(", "np.power("), {"x": 1.7, "y": 0.6, "np": np})
        assert abs(val - float(expr.subs({x: 1.7, y: 0.6}))) < 1e-12 * abs(val), code


MSH2 = """$MeshFormat
2.2 0 8
$EndMeshFormat
$Nodes
6
1 0 0 0
2 1 0 0
3 2 0 0
4 0 1 0
5 1 1 0
6 2 1 0
$EndNodes
$Elements
5
1 1 2 10 1 1 4
2 1 2 11 2 3 6
3 1 2 13 3 1 2
4 3 2 2 1 1 2 5 4
5 3 2 2 1 3 2 5 6
$EndElements
"""

MSH4 = """$MeshFormat
4.1 0 8
$EndMeshFormat
$Entities
0 2 1 0
1 0 0 0 0 1 0 1 10 0
2 2 0 0 2 1 0 1 11 0
1 0 0 0 2 1 0 1 2 0
$EndEntities
$Nodes
1 6 1 6
2 1 0 6
1
2
3
4
5
6
0 0 0
1 0 0
2 0 0
0 1 0
1 1 0
2 1 0
$EndNodes
$Elements
3 4 1 4
1 1 1 1
1 1 4
1 2 1 1
2 3 6
2 1 3 2
3 1 2 5 4
4 2 3 6 5
$EndElements
"""

GEO = """L = 2.; h = 1.;
Point(1) = {0, 0, 0, 1}; Point(2) = {L, 0, 0, 1}; Point(3) = {L, h, 0, 1}; Point(4) = {0, h, 0, 1};
Line(1) = {1, 2}; Line(2) = {2, 3}; Line(3) = {3, 4}; Line(4) = {4, 1};
Curve Loop(5) = {1, 2, 3, 4}; Plane Surface(1) = {5};
Physical Curve("in", 7) = {4}; Physical Curve("out", 8) = {2}; Physical Curve("wall", 9) = {1, 3};
Transfinite Curve{1, -3} = 5 Using Progression 2; Transfinite Curve{4, -2} = 4 Using Bump 0.1;
Transfinite Surface{1} = {1, 2, 3, 4}; Recombine Surface{1};
"""


def test_msh_reader_and_transfinite_geo(tmp_path):
    """`Mesh('x.msh')` (examples/bortels_threeion_nondim.py:111): gmsh 2.2 / 4.1 ASCII quads with physical boundary
    tags; a clockwise cell is re-oriented; without the .msh the structured mesh is built from the .geo."""
    from echemfem_b200.fd import Mesh
    from echemfem_b200.meshio import progression_nodes
    for name, text in (("a.msh", MSH2), ("b.msh", MSH4)):
        p = tmp_path / name
        p.write_text(text)
        m = Mesh(str(p))
        assert m.num_cells == 2 and m.dim == 2 and m.num_interior_facets == 1
        X = m.cell_coords()
        e1, e2 = X[:, 2] - X[:, 0], X[:, 1] - X[:, 0]
        assert np.all(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0] > 0)          # positively oriented local frames
        tags = sorted(int(t) for t in m.ext_marker.ravel() if t)
        assert tags == ([10, 11, 13] if name == "a.msh" else [10, 11])
    (tmp_path / "c.geo").write_text(GEO)
    m = Mesh(str(tmp_path / "c.msh"))
    assert m.shape == (4, 3)
    assert np.allclose(np.diff(m.axes[0]), 2.0 * np.array([1, 2, 4, 8]) / 15.0)
    assert np.allclose(m.axes[0], progression_nodes(2.0, 5, 2.0))
    assert np.isclose(m.axes[1][-1], 1.0) and np.allclose(m.axes[1], 1.0 - m.axes[1][::-1])   # Bump is symmetric
    assert set(np.unique(m.ext_marker)) == {0, 7, 8, 9}
    c0 = m.cell_coords()[0]
    assert np.allclose(c0[0], [0.0, 0.0])
    with pytest.raises(FileNotFoundError):
        Mesh(str(tmp_path / "missing.msh"))


@pytest.mark.parametrize("p,tile", [(1, None), (1, 4), (2, None)])
def test_multigrid_prolongation_embeds_the_coarse_dg_space(p, tile):
    """multigrid.Hierarchy: P maps the nodal values of a coarse piecewise polynomial to its nodal values on the
    fine mesh exactly (graded axes, tile-numbered cells, Q1 and Q2), so P^T A P is a Galerkin coarse operator."""
    from echemfem_b200.mesh import TensorMesh
    from echemfem_b200.multigrid import Hierarchy
    from echemfem_b200.function import FunctionSpace
    ax = [np.array([0.0, 0.1, 0.3, 0.35, 0.6, 0.7, 0.9, 0.95, 1.0]), np.linspace(0.0, 2.0, 9) ** 1.3]
    mesh = TensorMesh(ax, tile=tile)
    h = Hierarchy(mesh, p, tile=2, coarsest_cells=4)
    assert h.nlevels == 3 and [L.shape for L in h.levels] == [(8, 8), (4, 4), (2, 2)]
    Vf = FunctionSpace(mesh, "DG", p)
    xf = Vf.node_coords()
    # coarse level 1 as a mesh of its own, numbered as the hierarchy numbers it
    L1 = h.levels[1]
    mc = TensorMesh(L1.axes)
    if L1.perm is not None:
        from echemfem_b200.mesh import Mesh
        mc = Mesh(mc.coords, mc.cells[L1.perm], None)
    xc = FunctionSpace(mc, "DG", p).node_coords()
    n = (p + 1) ** 2
    # a different polynomial of degree p (per direction) in every coarse cell
    rng = np.random.default_rng(0)
    coef = rng.normal(size=(L1.ncell, 3))

    def f(x, cell):
        c = coef[cell]
        if p == 1:
            return c[:, 0] + c[:, 1] * x[:, 0] + c[:, 2] * x[:, 0] * x[:, 1]
        return c[:, 0] + c[:, 1] * x[:, 0] ** 2 * x[:, 1] + c[:, 2] * x[:, 1] ** 2
    cell_c = np.repeat(np.arange(L1.ncell), n)
    uc = f(xc, cell_c)
    # parent (in hierarchy numbering) of every fine dof: locate the fine node in the coarse axes
    ix = np.searchsorted(L1.axes[0], xf[:, 0], side="right") - 1
    iy = np.searchsorted(L1.axes[1], xf[:, 1], side="right") - 1
    lex = ix * L1.shape[1] + iy
    parent = lex if L1.inv is None else L1.inv[lex]
    uf = f(xf, parent)
    assert np.abs(h.P[0] @ uc - uf).max() < 1e-12 * np.abs(uf).max()
    # partition of unity: constants are reproduced on every level
    for P in h.P:
        assert np.allclose(P @ np.ones(P.shape[1]), 1.0)


def test_multigrid_hierarchy_3d_and_1d():
    """Hierarchy in one and three dimensions: level shapes, partition of unity, exact transfer of a coarse polynomial."""
    from echemfem_b200.mesh import TensorMesh
    from echemfem_b200.multigrid import Hierarchy
    from echemfem_b200.function import FunctionSpace
    m3 = TensorMesh([np.linspace(0, 1, 5), np.linspace(0, 2, 5) ** 1.2, np.linspace(0, 1, 9)], tile=(2, 2, 4))
    h = Hierarchy(m3, 1, tile=2, coarsest_cells=8)
    assert [L.shape for L in h.levels] == [(4, 4, 8), (2, 2, 4)]
    P = h.P[0]
    assert P.shape == (128 * 8, 16 * 8) and np.allclose(P @ np.ones(P.shape[1]), 1.0)
    # a globally trilinear function is in every level's space: nodal values map to nodal values
    xf = FunctionSpace(m3, "DG", 1).node_coords()
    L1 = h.levels[1]
    from echemfem_b200.mesh import Mesh
    mc = TensorMesh(L1.axes)
    if L1.perm is not None:
        mc = Mesh(mc.coords, mc.cells[L1.perm], None)
    xc = FunctionSpace(mc, "DG", 1).node_coords()
    f = lambda x: 1.0 + x[:, 0] - 2.0 * x[:, 1] * x[:, 2] + 0.5 * x[:, 0] * x[:, 1] * x[:, 2]
    assert np.abs(P @ f(xc) - f(xf)).max() < 1e-12
    m1 = TensorMesh([np.linspace(0, 1, 17) ** 2])
    h1 = Hierarchy(m1, 2, coarsest_cells=2)
    assert [L.shape for L in h1.levels] == [(16,), (8,), (4,), (2,)]
    for P in h1.P:
        assert np.allclose(P @ np.ones(P.shape[1]), 1.0)


def test_invalid_physics_is_rejected_at_construction():
    """tests/test_diffusion_finite_size.py:49-51 and tests/test_notimplemented.py of the reference (solver.py:115-142)."""
    import product_cases
    from echemfem_b200 import EchemSolver
    from echemfem_b200.fd import UnitIntervalMesh
    with pytest.raises(NotImplementedError):
        product_cases.DiffusionFiniteSizeSolver(2, family="DG")

    class Solver(EchemSolver):
        def __init__(self, flow):
            super().__init__([], {"flow": flow}, UnitIntervalMesh(2), family="CG")

        def set_boundary_markers(self):
            self.boundary_markers = {}
    for flow in (["electroneutrality"], ["poisson", "migration", "electroneutrality"],
                 ["poisson", "migration", "electroneutrality full"], ["migration"]):
        with pytest.raises(NotImplementedError):
            Solver(flow)
