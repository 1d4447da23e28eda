"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle on the CPU and the CUDA
path on the GPU against the same stored residuals / Jacobians of the five BASELINE.json configurations."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

NAMES = ["cfg1_bortels_2d", "cfg2_mms_dg_q1", "cfg3_gupta_cg_supg", "cfg4_bortels_3d", "cfg5_catalyst_layer"]


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


import parity  # noqa: E402




@pytest.fixture(scope="module")
def oracle_cases():
    return make_golden.cases()


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(oracle_cases, name):
    g = _load(name)
    d, u = oracle_cases[name]
    assert np.array_equal(u, g["u"])
    indptr, indices = d.pattern()
    assert np.array_equal(indptr, g["indptr"]) and np.array_equal(indices, g["indices"])
    parity.assert_vec_close(d.residual(u), g["F"], d.nf, name + " residual", 1e-13, scale=g["Fscale"])
    parity.assert_csr_close(d.jacobian(u).data, g["data"], indptr, indices, d.nf, d.ndof_scalar, d.n,
                            name + " jacobian", continuous=(d.prob.family != "DG"), rtol=1e-13, scale=g["Jscale"])


def _product(name):
    import product_cases as pc
    from echemfem_b200.mesh import ExtrudedMesh
    from echemfem_b200.meshes import BortelsMesh
    if name == "cfg2_mms_dg_q1":
        return pc.AdvectionDiffusionMigrationSolver(4, 1.0)
    if name == "cfg1_bortels_2d":
        return pc.BortelsSolver(BortelsMesh(3, 4, 3, 3))
    if name == "cfg3_gupta_cg_supg":
        return pc.GuptaSolver(4, 2, 2, SUPG=True)
    if name == "cfg4_bortels_3d":
        return pc.Bortels3DSolver(ExtrudedMesh(BortelsMesh(3, 3, 3, 3), 2, 3.0), 6.0)
    return pc.CatalystLayerSolver(4, p=1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_golden(name):
    import torch
    from echemfem_b200.newton import NewtonEngine
    g = _load(name)
    s = _product(name)
    eng = NewtonEngine(s)
    s.u.tensor.copy_(torch.from_numpy(g["u"]).cuda())
    rp, ci, _ = eng.csr_host()
    assert np.array_equal(rp, g["indptr"]) and np.array_equal(ci, g["indices"])          # sparsity: bit-exact
    F = torch.zeros_like(s.u.tensor)
    eng.residual_jacobian(s.u.tensor, out=F)
    parity.assert_vec_close(F.cpu().numpy(), g["F"], eng.nf, name + " residual", scale=g["Fscale"])
    parity.assert_csr_close(eng.vals.cpu().numpy(), g["data"], g["indptr"], g["indices"], eng.nf, eng.N, eng.nloc,
                            name + " jacobian", continuous=(eng.family == "CG"), scale=g["Jscale"])
