"""Entry-level parity beyond toy meshes: the CUDA path through the host-buffer C ABI (`ech_assemble_host`, 16-chunk
pipeline) against the C++ CPU port (oracle/cpu_port, itself checked against the numpy oracle in
tests/test_cpu_port.py) on the BENCHMARK physics (cfg2, D = (0.5e-5, 1e-5)):

* a medium mesh (259^2 = 67 081 cells: ragged last warp batch, ragged last host chunk), every entry;
* the full benchmark mesh (1408^2, 15.86 M dofs, 634 M non-zeros): the rows of three cell ranges (first, middle,
  ragged end), so that int64 row pointers, uint8 slot tables and the chunked pipeline are CHECKED at the size the
  headline number is quoted on, not only timed there.

Tolerance: tests/parity.py (1e-12 * tau_block with the port's element-level scales)."""
import numpy as np
import pytest
import torch

import parity
from echemfem_b200 import capi
from echemfem_b200.capi import ptr
from echemfem_b200.workloads import cfg2_solver, cfg2_state

pytestmark = pytest.mark.gpu


def _host_assemble(eng, s, u_host, chunks):
    lib = eng.lib
    capi.check(lib.ech_set_host_chunks(eng.h, chunks), eng.h)
    uh = torch.from_numpy(u_host.copy()).pin_memory()
    Fh = torch.full((u_host.size,), np.nan, dtype=torch.float64).pin_memory()
    eng.vals.fill_(float("nan"))
    aux, consts = eng._aux_tensor(), eng._consts()
    capi.check(lib.ech_assemble_host(eng.h, ptr(uh), ptr(s.u.tensor), ptr(aux), ptr(consts), ptr(eng.rowptr),
                                     ptr(eng.vals), ptr(eng.F), ptr(Fh), eng._stream()), eng.h)
    torch.cuda.synchronize()
    return Fh.numpy()


@pytest.mark.parametrize("p,N", [(1, 259), (2, 61)])
def test_medium_mesh_every_entry_matches_cpu_port(p, N):
    from oracle.cpu_port import CpuPort
    from echemfem_b200.newton import NewtonEngine
    s = cfg2_solver(N, p=p)
    eng = NewtonEngine(s)
    u = cfg2_state(s.V.node_coords())
    port = CpuPort(s)
    rp = eng.rowptr.cpu().numpy()
    assert np.array_equal(rp, port.rowptr)                       # sparsity: row pointers bit-exact
    Fp, Jp, Fs, Js = port.assemble(u, scales=True)
    F = _host_assemble(eng, s, u, 16)
    vals = eng.vals.cpu().numpy()
    assert np.isfinite(vals).all() and np.isfinite(F).all()
    ci = eng.colidx.cpu().numpy()
    parity.assert_vec_close(F, Fp, eng.nf, "Q%d N=%d residual" % (p, N), scale=Fs.reshape(eng.nf, -1).max(axis=1))
    parity.assert_csr_close(vals, Jp, rp, ci, eng.nf, eng.N, eng.nloc, "Q%d N=%d jacobian" % (p, N), scale=Js)
    # unchunked call and device-resident call: bit-identical
    F1 = _host_assemble(eng, s, u, 0)
    assert np.array_equal(F1, F) and np.array_equal(eng.vals.cpu().numpy(), vals)


def test_full_benchmark_mesh_sampled_rows_match_cpu_port():
    from oracle.cpu_port import CpuPort
    from echemfem_b200.newton import NewtonEngine
    s = cfg2_solver()                                            # 1408^2, the mesh BENCH is quoted on
    eng = NewtonEngine(s)
    assert eng.ndof == 15859712 and eng.nnz > 2 ** 29
    u = cfg2_state(s.V.node_coords())
    F = _host_assemble(eng, s, u, 16)
    port = CpuPort(s)
    rp = eng.rowptr.cpu().numpy()
    assert np.array_equal(rp, port.rowptr)
    nc, n, nf, N = eng.ncell, eng.nloc, eng.nf, eng.N
    for c0, c1 in ((0, 4099), (nc // 2 - 2048, nc // 2 + 2051), (nc - 4093, nc)):
        Fp = np.zeros(port.ndof)
        Jp = np.zeros(port.nnz)
        _, _, Fs, Js = port.assemble(u, cells=(c0, c1), scales=True, out=(Fp, Jp))
        for f in range(nf):
            r0, r1 = f * N + c0 * n, f * N + c1 * n
            e0, e1 = int(rp[r0]), int(rp[r1])
            got = eng.vals[e0:e1].cpu().numpy()
            ci = eng.colidx[e0:e1].cpu().numpy()
            parity.assert_csr_close(got, Jp[e0:e1], rp[r0:r1 + 1] - e0, ci, nf, N, n,
                                    "cells [%d, %d) field %d jacobian" % (c0, c1, f), scale=Js[e0:e1])
            tau = Fs[r0:r1].max()
            assert np.abs(F[r0:r1] - Fp[r0:r1]).max() <= 1e-12 * tau, (c0, c1, f)
