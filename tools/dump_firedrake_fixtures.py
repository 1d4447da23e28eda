#!/usr/bin/env python
"""Dump reference fixtures on a host that HAS Firedrake + the reference echemfem installed (last row).

Not runnable in the build container (Firedrake / PETSc are not installable offline); written against
the Firedrake API the reference itself uses (`NonlinearVariationalProblem`, `assemble`, `derivative`,
`cell_node_map`, `interior_facets`), to be run ONCE by a maintainer:

    python tools/dump_firedrake_fixtures.py --out tests/golden/firedrake

For every configuration it writes `<name>.npz` with everything needed to pin the conventions the oracle restates
 and to compare entry by entry WITHOUT reproducing Firedrake's dof numbering:

    u                 state the forms are assembled at (mixed vector, Firedrake's ordering)
    F                 assemble(Form) at u
    J_indptr/J_indices/J_data     assemble(derivative(Form, u)) as CSR (PETSc AIJ of the mixed space, bcs NOT applied)
    dof_coords_<i>    coordinates of the dofs of sub-space i (matching key for DG: (cell, local node) via cell_node_map;
                      for CG: coordinates)
    cell_node_map_<i> V_i.cell_node_map().values
    cell_vertices     coordinates of the vertices of every cell in Firedrake's local vertex order
    int_facet_cells / int_facet_local / ext_facet_cells / ext_facet_local / ext_facet_markers
    quad_points/quad_weights/basis_at_quad   the cell rule of degree p+1 and the tabulated scalar basis (variant pin)
    field_offsets     start of every sub-space in the mixed vector

`tests/test_golden.py` can then load these next to the oracle's own fixtures; the matching is by (field, cell, local
node) for DG and by coordinates for CG, as SURVEY 8c Q7 prescribes.
"""
import argparse
import os

import numpy as np


def dump(name, solver, out_dir):
    import firedrake as fd
    from firedrake import assemble, derivative
    from firedrake.petsc import PETSc  # noqa: F401
    W, u, Form = solver.W, solver.u, solver.Form
    mesh = solver.mesh
    F = assemble(Form)
    J = assemble(derivative(Form, u), mat_type="aij")
    indptr, indices, data = J.petscmat.getValuesCSR()
    rec = {"u": np.concatenate([s.dat.data_ro.reshape(-1) for s in u.subfunctions]),
           "F": np.concatenate([s.dat.data_ro.reshape(-1) for s in F.subfunctions]),
           "J_indptr": indptr, "J_indices": indices, "J_data": data}
    off = [0]
    for i, V in enumerate(W.subfunctions if hasattr(W, "subfunctions") else W):
        Vv = fd.VectorFunctionSpace(mesh, V.ufl_element())
        xs = fd.Function(Vv).interpolate(fd.SpatialCoordinate(mesh))
        rec["dof_coords_%d" % i] = xs.dat.data_ro.copy()
        rec["cell_node_map_%d" % i] = V.cell_node_map().values.copy()
        off.append(off[-1] + V.dof_dset.size)
    rec["field_offsets"] = np.array(off)
    coords = mesh.coordinates
    rec["cell_vertices"] = coords.dat.data_ro[coords.cell_node_map().values].copy()
    IF, EF = mesh.interior_facets, mesh.exterior_facets
    rec["int_facet_cells"] = IF.facet_cell.copy()
    rec["int_facet_local"] = IF.local_facet_dat.data_ro.copy()
    rec["ext_facet_cells"] = EF.facet_cell.copy()
    rec["ext_facet_local"] = EF.local_facet_dat.data_ro.copy()
    rec["ext_facet_markers"] = np.asarray(EF.markers)
    # quadrature rule of degree p+1 and the tabulated scalar basis of sub-space 0 (nodal variant, local ordering)
    from finat.quadrature import make_quadrature
    from tsfc.finatinterface import create_element
    V0 = (W.subfunctions if hasattr(W, "subfunctions") else W)[0]
    el = create_element(V0.ufl_element())
    rule = make_quadrature(el.cell, solver.poly_degree + 1)
    rec["quad_points"] = np.asarray(rule.point_set.points)
    rec["quad_weights"] = np.asarray(rule.weights)
    tab = el.fiat_equivalent.tabulate(1, rule.point_set.points)
    for k, v in tab.items():
        rec["basis_at_quad_%s" % "".join(str(i) for i in k)] = np.asarray(v)
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
    print("wrote", name, "ndof", len(rec["u"]), "nnz", len(data))


def perturbed(solver, amp=1e-3):
    """The benchmark state of SURVEY 8d: current state + amp * sin(7x + 3y) on every field."""
    import firedrake as fd
    x = fd.SpatialCoordinate(solver.mesh)
    for s in solver.u.subfunctions:
        s.interpolate(s + amp * fd.sin(7 * x[0] + 3 * x[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tests/golden/firedrake")
    ap.add_argument("--reference-tests", default=None, help="path of the reference's tests/ directory")
    args = ap.parse_args()
    import importlib.util
    import sys
    ref = args.reference_tests or os.path.join(os.path.dirname(__import__("echemfem").__file__), "..", "tests")

    def load(fname):
        spec = importlib.util.spec_from_file_location("ref_" + fname[:-3], os.path.join(ref, fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        return mod
    # configs[1] (cfg2 physics at toy size, Q1..Q3) and the CG / Poisson / porous twins of the reference's own tests
    adm = load("test_advection_diffusion_migration.py")
    for p in (1, 2, 3):
        s = adm.AdvectionDiffusionMigrationSolver(4, 1.0, p=p) if "p" in adm.AdvectionDiffusionMigrationSolver.__init__.__code__.co_varnames \
            else adm.AdvectionDiffusionMigrationSolver(4, 1.0)
        s.setup_solver(initial_solve=False)
        perturbed(s)
        dump("adm_dg_q%d_N4" % p, s, args.out)
        if "p" not in adm.AdvectionDiffusionMigrationSolver.__init__.__code__.co_varnames:
            break
    for fname, cls, a, tag in (("test_diffusion_migration_poisson.py", "DiffusionMigrationSolver", (4,), "poisson_N4"),
                               ("test_advection_diffusion_migration_porous.py", "AdvectionDiffusionMigrationSolver", (4, 1.0), "porous_N4"),
                               ("test_SUPG.py", "AdvectionDiffusionSolver", (4, 1e-3), "supg_cg_N4")):
        try:
            m = load(fname)
            s = getattr(m, cls)(*a)
            s.setup_solver(initial_solve=False)
            perturbed(s)
            dump(tag, s, args.out)
        except Exception as e:                       # a class name that differs between versions must not stop the rest
            print("skipped", fname, "->", repr(e))


if __name__ == "__main__":
    main()
