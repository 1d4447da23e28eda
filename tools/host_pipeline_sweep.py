"""Sweep of the host-buffer pipeline of ech_assemble_host on the cfg2 mesh: number of cell ranges.
usage: python tools/host_pipeline_sweep.py [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import product_cases
from echemfem_b200 import capi
from echemfem_b200.capi import ptr
from echemfem_b200.newton import NewtonEngine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1408
s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, D1=0.5e-5, D2=1.0e-5)
eng = NewtonEngine(s)
x = s.V.node_coords()
u_host = np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3, np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3])
u = s.u.tensor
uh = torch.from_numpy(u_host).pin_memory()
Fh = torch.empty(u_host.size, dtype=torch.float64).pin_memory()
aux, consts = eng._aux_tensor(), eng._consts()
lib, st = eng.lib, eng._stream()


def call():
    capi.check(lib.ech_assemble_host(eng.h, ptr(uh), ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr), ptr(eng.vals),
                                     ptr(eng.F), ptr(Fh), st), eng.h)


def timed(steps=20):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


capi.check(lib.ech_set_host_chunks(eng.h, 0), eng.h)
print("unchunked %.3f ms" % timed())
for chunks in (4, 8, 12, 16, 24, 32, 48, 64, 96):
    capi.check(lib.ech_set_host_chunks(eng.h, chunks), eng.h)
    print("chunks %2d  %.3f ms" % (chunks, timed()), flush=True)
