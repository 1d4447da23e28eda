"""Product weak form (symbolic restatement) vs oracle weak form (numpy restatement), point by point.

Two independent restatements of echemfem/solver.py:1461-1733, 1866-2013 must
give the same f0/f1 coefficients at random point data; tolerance 1e-12 relative
(the entry tolerance of BASELINE.json's north_star).
"""
import numpy as np
import pytest

from echemfem_b200.compiler import CompiledForm
from pointwise import random_point_data, evaluate
import mms_cases
import product_cases

RTOL = 1e-12


def _close(a, b):
    scale = max(np.abs(a).max(), np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= RTOL * scale * 50


@pytest.mark.parametrize("vavg", [1.0, 1e5])
def test_adm_dg_pointwise(vavg):
    s = product_cases.AdvectionDiffusionMigrationSolver(2, vavg)
    cf = CompiledForm(s.Form, 2, len(s._aux), 2, 1, "DG")
    prob = mms_cases.adm_problem(vavg)
    rng = np.random.default_rng(1)
    npts = 64
    dat = random_point_data(2, 1, 2, npts, rng)
    x = dat["x"][None]
    n = dat["n"][None]
    u = [dat["u"][j][None] for j in range(2)]
    gu = [dat["gu"][j][None] for j in range(2)]
    um = [dat["um"][j][None] for j in range(2)]
    gum = [dat["gum"][j][None] for j in range(2)]
    # cell
    f0, f1 = prob.cell(x, u, gu, {"cell_volume": dat["cv"][None], "cell_diameter": dat["cd"][None]})
    for i in range(2):
        assert _close(evaluate(cf.cell["f0"][i], dat, cf), f0[i][0])
        for k in range(2):
            assert _close(evaluate(cf.cell["f1"][i][k], dat, cf), f1[i][0, :, k])
    # interior facet, '+' side test functions
    geop = {"cell_volume": dat["cv"][None], "facet_area": dat["fa"][None]}
    geom = {"cell_volume": dat["cvm"][None], "facet_area": dat["fa"][None]}
    f0p, f1p, f0m, f1m = prob.interior_facet(x, n, u, gu, um, gum, geop, geom)
    for i in range(2):
        assert _close(evaluate(cf.int["f0"][i], dat, cf), f0p[i][0])
        for k in range(2):
            assert _close(evaluate(cf.int["f1"][i][k], dat, cf), f1p[i][0, :, k])
    # +/- swap symmetry: the '-' rows are the '+' rows of the swapped facet
    sw = dict(dat)
    sw.update({"u": dat["um"], "um": dat["u"], "gu": dat["gum"], "gum": dat["gu"],
               "cv": dat["cvm"], "cvm": dat["cv"], "n": -dat["n"], "a": dat["am"], "am": dat["a"]})
    for i in range(2):
        assert _close(evaluate(cf.int["f0"][i], sw, cf), f0m[i][0])
        for k in range(2):
            assert _close(evaluate(cf.int["f1"][i][k], sw, cf), f1m[i][0, :, k])
    # exterior facets, every marker
    for marker in (1, 2, 3, 4):
        cls = cf.marker_class[marker]
        g0, g1 = prob.exterior_facet(marker, x, n, u, gu, geop, {"U_app": dat["a"][0][None]})
        for i in range(2):
            assert _close(evaluate(cf.ext[cls]["f0"][i], dat, cf), g0[i][0])
            for k in range(2):
                assert _close(evaluate(cf.ext[cls]["f1"][i][k], dat, cf), g1[i][0, :, k])


def test_block_masks_match_oracle():
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    s = product_cases.AdvectionDiffusionMigrationSolver(2, 1.0)
    cf = CompiledForm(s.Form, 2, len(s._aux), 2, 1, "DG")
    disc = Discretization(unit_square_mesh(3), mms_cases.adm_problem(1.0))
    own, nbr = disc.block_mask()
    assert np.array_equal(own, cf.own_mask)
    assert np.array_equal(nbr, cf.nbr_mask)


def test_porous_pointwise():
    """Porous model (two potentials, Bruggeman coefficients): product vs oracle forms at random points."""
    s = product_cases.PorousSolver(2, 2.0)
    cf = CompiledForm(s.Form, 3, len(s._aux), 2, 1, "DG")
    prob = mms_cases.porous_problem(2.0)
    rng = np.random.default_rng(2)
    dat = random_point_data(3, 1, 2, 48, rng)
    x, n = dat["x"][None], dat["n"][None]
    u = [dat["u"][j][None] for j in range(3)]
    gu = [dat["gu"][j][None] for j in range(3)]
    um = [dat["um"][j][None] for j in range(3)]
    gum = [dat["gum"][j][None] for j in range(3)]
    f0, f1 = prob.cell(x, u, gu, {"cell_volume": dat["cv"][None], "cell_diameter": dat["cd"][None]})
    geop = {"cell_volume": dat["cv"][None], "facet_area": dat["fa"][None]}
    geom = {"cell_volume": dat["cvm"][None], "facet_area": dat["fa"][None]}
    f0p, f1p, _, _ = prob.interior_facet(x, n, u, gu, um, gum, geop, geom)
    g0, g1 = prob.exterior_facet(1, x, n, u, gu, geop, {"U_app": dat["a"][0][None]})
    cls = cf.marker_class[1]
    for i in range(3):
        assert _close(evaluate(cf.cell["f0"][i], dat, cf), f0[i][0])
        assert _close(evaluate(cf.int["f0"][i], dat, cf), f0p[i][0])
        assert _close(evaluate(cf.ext[cls]["f0"][i], dat, cf), g0[i][0])
        for k in range(2):
            assert _close(evaluate(cf.cell["f1"][i][k], dat, cf), f1[i][0, :, k])
            assert _close(evaluate(cf.int["f1"][i][k], dat, cf), f1p[i][0, :, k])
            assert _close(evaluate(cf.ext[cls]["f1"][i][k], dat, cf), g1[i][0, :, k])


@pytest.mark.parametrize("kind", ["BMCSL", "BK", "CS", "SP", "SMPNP", "finite size"])
def test_finite_size_and_mass_action_pointwise(kind):
    """Finite-size chemical potentials (solver.py:1543-1636) and law-of-mass-action reactions
    (solver.py:1222-1274): product (sympy chain rule) vs oracle (forward-mode AD) at random points."""
    s = product_cases.FiniteSizeSolver(2, kind)
    cf = CompiledForm(s.Form, 4, len(s._aux), 2, 1, "CG")
    prob = mms_cases.finite_size_problem(kind)
    rng = np.random.default_rng(5)
    dat = random_point_data(4, 1, 2, 32, rng)
    x = dat["x"][None]
    u = [dat["u"][j][None] for j in range(4)]
    gu = [dat["gu"][j][None] for j in range(4)]
    f0, f1 = prob.cell(x, u, gu, {"cell_volume": dat["cv"][None], "cell_diameter": dat["cd"][None]})
    for i in range(4):
        assert _close(evaluate(cf.cell["f0"][i], dat, cf), f0[i][0]), (kind, i)
        for k in range(2):
            assert _close(evaluate(cf.cell["f1"][i][k], dat, cf), f1[i][0, :, k]), (kind, i, k)


def test_finite_size_rejected_on_dg():
    from oracle.forms import Problem
    with pytest.raises(NotImplementedError):
        Problem([{"name": "A", "diffusion coefficient": 1.0, "z": 0.0}],
                {"flow": ["diffusion", "finite size"], "Avogadro constant": 1.0}, {}, family="DG")
