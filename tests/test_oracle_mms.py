"""The oracle against the reference's own acceptance criteria (its only "golden" data).

Restates tests/test_advection_diffusion_migration.py:66-95 of the reference: the L2 errors of
C1 and U must drop by the stated ratios per refinement of UnitSquareMesh(N, N) (DG Q1).
The reference ships no golden vectors; these inequalities are what pins the oracle.
"""
import numpy as np
import pytest

import mms_cases
from oracle.mesh import unit_square_mesh
from oracle.assemble import Discretization
from oracle.newton import solve_problem


def _errors(N, vavg, p=1):
    disc = Discretization(unit_square_mesh(N), mms_cases.adm_problem(vavg, p=p))
    u, info = solve_problem(disc)
    Ns = disc.ndof_scalar
    return disc.l2_error(u[:Ns], mms_cases.C1ex), disc.l2_error(u[Ns:], mms_cases.Uex), info


def test_convergence_low_peclet():
    eC_old, eU_old = 1e6, 1e6
    for i in range(4):
        eC, eU, info = _errors(2 ** (i + 1), 1.0)
        assert eC < 0.26 * eC_old          # reference test: tests/test_advection_diffusion_migration.py:75
        assert eU < 0.26 * eU_old
        eC_old, eU_old = eC, eU


def test_convergence_high_peclet():
    eC_old, eU_old = 1e6, 1e6
    for i in range(4):
        eC, eU, info = _errors(2 ** (i + 1), 1e5)
        assert eC < 0.36 * eC_old          # p + 1/2 convergence, reference test :92
        assert eU < 0.26 * eU_old
        eC_old, eU_old = eC, eU


def test_jacobian_is_derivative_of_residual():
    """Complex-step Jacobian vs central finite differences of the oracle residual."""
    disc = Discretization(unit_square_mesh(3), mms_cases.adm_problem(3.0))
    rng = np.random.default_rng(0)
    x = disc.node_coords
    u = np.concatenate([mms_cases.C1ex(x), mms_cases.Uex(x)]) + 1e-2 * rng.standard_normal(disc.ndof)
    J = disc.jacobian(u)
    v = rng.standard_normal(disc.ndof)
    h = 1e-6
    fd = (disc.residual(u + h * v) - disc.residual(u - h * v)) / (2 * h)
    assert np.linalg.norm(J @ v - fd) <= 1e-7 * np.linalg.norm(fd)


def test_pattern_structure():
    """Field-blocked AIJ pattern: sorted columns, diagonal present, nnz = nf^2 n^2 (Ncell + 2 Nint)."""
    N = 4
    disc = Discretization(unit_square_mesh(N), mms_cases.adm_problem(1.0))
    indptr, indices = disc.pattern()
    nint = disc.mesh.int_facets.shape[0]
    assert indptr[-1] == 4 * 16 * (N * N + 2 * nint)
    for r in range(disc.ndof):
        cols = indices[indptr[r]:indptr[r + 1]]
        assert np.all(np.diff(cols) > 0)
        assert r in cols
