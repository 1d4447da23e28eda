"""The analysis behind the multigrid preconditioner, on the CPU (tests/mg_prototype.py: scipy cycle on oracle
Jacobians of the cfg2 physics).  Pins the qualitative claims of echemfem_b200/auxspace.py at a size that runs in
seconds: the conforming auxiliary space for the potential does not hurt at small N and removes the growth of the
iteration count with N; for the advection-dominated concentration field it destroys the cycle."""
import numpy as np
import pytest

import mg_prototype as mp


def test_transfer_operators_are_exact_embeddings():
    f, c = mp.make_disc(8), mp.make_disc(4)
    g = lambda x: 2 * x[:, 0] - 3 * x[:, 1] + 0.5 * x[:, 0] * x[:, 1]
    P = mp.dg_prolongation(f, 8, c, 4)
    assert np.abs(P @ g(c.node_coords) - g(f.node_coords)).max() < 1e-13
    Q = mp.cg_embedding(f, 8)
    gx, gy = np.meshgrid(np.linspace(0, 1, 9), np.linspace(0, 1, 9), indexing="ij")
    V = np.stack([gx.ravel(), gy.ravel()], axis=1)
    assert np.abs(Q @ g(V) - g(f.node_coords)).max() < 1e-13
    # a continuous function has no jumps: neighbouring cells see the same vertex value
    assert Q.shape == (f.ndof_scalar, 81) and abs(Q.sum(axis=1) - 1).max() < 1e-14


def test_auxiliary_space_flattens_the_iteration_growth():
    plain = {N: mp.iterations(N) for N in (32, 64)}
    aux = {N: mp.iterations(N, (1,)) for N in (32, 64)}
    # measured: 23 / 30 without, 25 / 27 with (N = 128: 45 / 32)
    assert plain[64] >= plain[32] + 4
    assert aux[64] <= aux[32] + 3
    assert aux[64] < plain[64]
    print("prototype iterations N = 32 / 64: nested DG hierarchy %s, with the conforming space for the potential %s"
          % (list(plain.values()), list(aux.values())))


def test_conforming_space_for_the_concentration_field_breaks_the_cycle():
    assert mp.iterations(32, (0, 1), maxit=120) == 120


@pytest.mark.parametrize("ranks,R", [(2, 1), (2, 2)])
def test_global_conforming_space_removes_the_penalty_of_the_rank_cut(ranks, R):
    """Strips along x (R = 1: the unit square cut; R = ranks: the weak-scaling layout, one N x N strip per rank):
    block Jacobi of the rank-local cycles needs several times the iterations of one rank on the same domain; the
    conforming space on the global vertex grid (auxspace.GlobalConformingCoarse) brings them back
    (N = 64: 27 / 133 / 26 and 30 / 136 / 29)."""
    one, bj, shared = mp.two_rank_iterations(32, ranks=ranks, R=R)
    print("prototype, %d strips of [0, %d] x [0, 1] at N = 32: one rank %d, block Jacobi %d, + shared conforming space %d"
          % (ranks, R, one, bj, shared))
    assert bj >= 2 * one
    assert shared <= one + 3
