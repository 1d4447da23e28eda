"""The C++ CPU port (oracle/cpu_port: bench.py's CPU baseline and the large-mesh oracle) against the numpy oracle.

Both are test infrastructure; the port shares only the generated pointwise integrands with the product, the numpy
oracle shares nothing.  Entry tolerance: tests/parity.py (1e-12 * tau_block)."""
import numpy as np
import pytest

import mms_cases
import parity
import product_cases


def _check(disc, s, u, what):
    from oracle.cpu_port import CpuPort
    port = CpuPort(s)
    indptr, indices = disc.pattern()
    assert np.array_equal(port.rowptr, indptr), what
    F, vals = port.assemble(u)
    parity.assert_vec_close(F, disc.residual(u), port.nf, what + " residual", scale=disc.residual_scale(u))
    J, Jscale = disc.jacobian(u, return_scale=True)
    parity.assert_csr_close(vals, J.data, indptr, indices, port.nf, port.N, port.n, what + " jacobian", scale=Jscale)
    # the port's element-level scales are the oracle's (same pieces: cell integral + facet halves)
    _, _, Fs, vs = port.assemble(u, scales=True)
    assert np.abs(vs - Jscale).max() <= 1e-10 * Jscale.max(), what
    tau = disc.residual_scale(u)          # also counts analytically cancelling pointwise sums, so it may be larger
    for f in range(port.nf):
        fs = Fs.reshape(port.nf, -1)[f]
        assert np.all(fs >= np.abs(F.reshape(port.nf, -1)[f]) * (1 - 1e-12)) and fs.max() <= 8 * tau[f], what
    # cell ranges (threads / chunks) write disjoint rows
    F2, vals2 = port.assemble(u, cells=(0, port.ncell // 2))
    F3, vals3 = port.assemble(u, cells=(port.ncell // 2, port.ncell))
    assert np.array_equal(F2 + F3, F) and np.array_equal(vals2 + vals3, vals)


def _mms_state(x):
    ph = 7 * x[:, 0] + 3 * x[:, 1] + (5 * x[:, 2] if x.shape[1] > 2 else 0.0)
    pert = 1e-3 * np.sin(ph)
    return np.concatenate([mms_cases.C1ex(x) + pert, mms_cases.Uex(x) - pert])


@pytest.mark.parametrize("p,N,D", [(1, 5, (0.5, 1.0)), (1, 6, (0.5e-5, 1.0e-5)), (2, 3, (0.5e-5, 1.0e-5)), (3, 2, (0.5, 1.0))])
def test_port_matches_numpy_oracle_mms(p, N, D):
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    disc = Discretization(unit_square_mesh(N), mms_cases.adm_problem(1.0, D1=D[0], D2=D[1], p=p))
    s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, D1=D[0], D2=D[1], p=p)
    _check(disc, s, _mms_state(disc.node_coords), "Q%d" % p)


def test_port_matches_numpy_oracle_bortels():
    """Graded mesh, three fields (7 of 9 blocks), Butler-Volmer electrodes, inlet / outlet."""
    from oracle.assemble import Discretization
    from echemfem_b200.meshes import BortelsMesh
    n = (4, 5, 3, 4)
    disc = Discretization(mms_cases.bortels_oracle_mesh(*n), mms_cases.bortels_problem())
    s = product_cases.BortelsSolver(BortelsMesh(*n))
    x = disc.node_coords
    pert = 1e-2 * np.sin(3 * x[:, 0] + 5 * x[:, 1])
    _check(disc, s, np.concatenate([1.0 + pert, 1.0 - 0.5 * pert, 0.1 + pert]), "bortels")


def test_port_matches_numpy_oracle_hexes():
    from oracle.mesh import unit_cube_mesh
    from oracle.assemble import Discretization
    from echemfem_b200.mesh import UnitSquareMesh, ExtrudedMesh
    disc = Discretization(unit_cube_mesh(3, 2, 2), mms_cases.adm_problem(1.0))
    s = product_cases.AdvectionDiffusionMigrationSolver(0, 1.0, mesh=ExtrudedMesh(UnitSquareMesh(3, 2), 2, 0.5))
    _check(disc, s, _mms_state(disc.node_coords), "hex")
