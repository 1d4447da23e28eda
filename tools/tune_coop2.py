"""Time the fused DG kernel of the cfg2 module (device-resident, CUDA events) for the current build / environment.
usage: [ECH_COOP2_WAVES=w] [ECH_MODULE_FLAGS="-D..."] python tools/tune_coop2.py [N] [degree]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from echemfem_b200.workloads import cfg2_solver, cfg2_state
from echemfem_b200.newton import NewtonEngine
from echemfem_b200 import capi
from echemfem_b200.capi import ptr

deg = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N = int(sys.argv[1]) if len(sys.argv) > 1 else {1: 1408, 2: 944, 3: 704}[deg]
sys.stdout = sys.stderr
s = cfg2_solver(N, p=deg)
eng = NewtonEngine(s)
u = s.u.tensor
u.copy_(torch.from_numpy(cfg2_state(s.V.node_coords())).cuda())
aux, consts = eng._aux_tensor(), eng._consts()
lib, st = eng.lib, eng._stream()


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


tfj = timed(lambda: capi.check(lib.ech_residual_jacobian(eng.h, ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr), ptr(eng.vals), ptr(eng.F), st), eng.h))
tj = timed(lambda: capi.check(lib.ech_jacobian(eng.h, ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr), ptr(eng.vals), st), eng.h))
tf = timed(lambda: capi.check(lib.ech_residual(eng.h, ptr(u), ptr(aux), ptr(consts), ptr(eng.F), st), eng.h))
print("TUNE waves=%s flags=%s N=%d Q%d: fused %.3f ms  J %.3f ms  F %.3f ms  (%.0f MDOF/s fused)" % (
    os.environ.get("ECH_COOP2_WAVES", "default"), os.environ.get("ECH_MODULE_FLAGS", ""), N, deg, tfj, tj, tf,
    eng.ndof / tfj / 1e3), file=sys.__stdout__)
