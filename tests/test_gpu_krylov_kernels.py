"""Vector kernels of the Krylov loop against numpy: fused multi-dot, multi-axpy and the fused update + projection of
the Gram-Schmidt refinement pass (`ech_multidot`, `ech_multiaxpy`, `ech_multiaxpy_dot`; PETSc's VecMDot / VecMAXPY
inside KSPGMRESClassicalGramSchmidtOrthogonalization, which the reference's solves run through)."""
import numpy as np
import pytest
import torch

from echemfem_b200 import capi
from echemfem_b200.capi import ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 3), (100003, 16), (70001, 17), (50021, 33), (40009, 64)])
def test_multiaxpy_dot_matches_numpy(n, k):
    lib = capi.lib()
    rng = np.random.default_rng(n + k)
    V = rng.standard_normal((k, n))
    w = rng.standard_normal(n)
    h = rng.standard_normal(k)
    Vd, wd, hd = (torch.from_numpy(a).cuda() for a in (V.reshape(-1), w, h))
    work = torch.empty(4096 + (k + 2) * 592, dtype=torch.float64, device="cuda")
    out = torch.empty(k + 1, dtype=torch.float64, device="cuda")
    st = capi.C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # the two separate kernels
    w1 = wd.clone()
    capi.check(lib.ech_multiaxpy(n, k, ptr(Vd), ptr(hd), -1.0, ptr(w1), st))
    capi.check(lib.ech_multidot(n, k, ptr(Vd), ptr(w1), ptr(work), ptr(out), st))
    ref_w = w - V.T @ h
    ref = np.concatenate([V @ ref_w, [ref_w @ ref_w]])
    scale = np.abs(V).max() * np.abs(ref_w).max() * n
    np.testing.assert_allclose(w1.cpu().numpy(), ref_w, rtol=0, atol=1e-13 * (1 + np.abs(h).sum() * np.abs(V).max()))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-14 * scale)
    # fused; h and out in the same buffer
    w2 = wd.clone()
    buf = torch.zeros(k + 1, dtype=torch.float64, device="cuda")
    buf[:k] = hd
    capi.check(lib.ech_multiaxpy_dot(n, k, ptr(Vd), ptr(buf), -1.0, ptr(w2), ptr(work), ptr(buf), st))
    np.testing.assert_allclose(w2.cpu().numpy(), w1.cpu().numpy(), rtol=0, atol=1e-14 * (1 + np.abs(ref_w).max()))
    np.testing.assert_allclose(buf.cpu().numpy(), ref, rtol=0, atol=1e-14 * scale)


def test_multiaxpy_dot_rejects_long_bases():
    lib = capi.lib()
    t = torch.zeros(8, dtype=torch.float64, device="cuda")
    st = capi.C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.ech_multiaxpy_dot(8, 65, ptr(t), ptr(t), 1.0, ptr(t), ptr(t), ptr(t), st) != 0
