"""CG assembly timing on the configs[2] physics (Gupta, five fields, CG Q1) at increasing mesh sizes (GPU box).
Not the headline bench: a record of where the node-owner-gather CG kernels stand."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import product_cases
from echemfem_b200.newton import NewtonEngine

sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(400, 200)]
for nx, ny in sizes:
    for supg in (False, True):
        t0 = time.time()
        s = product_cases.GuptaSolver(nx, ny, ny, SUPG=supg)
        eng = NewtonEngine(s)
        t1 = time.time()
        x = s.V.node_coords()
        ph = np.sin(7e2 * x[:, 0] + 3e3 * x[:, 1])
        bulk = [b for _, _, b, _ in product_cases.GUPTA_SPECIES[:4]]
        u = np.concatenate([b * (1.0 + 0.05 * ph) for b in bulk] + [1e-3 * ph])
        s.u.tensor.copy_(torch.from_numpy(u).cuda())
        for _ in range(2):
            eng.residual(s.u.tensor)
            eng.jacobian(s.u.tensor)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 5
        ev[0].record()
        for _ in range(reps):
            eng.residual(s.u.tensor)
        ev[1].record()
        for _ in range(reps):
            eng.jacobian(s.u.tensor)
        ev[2].record()
        torch.cuda.synchronize()
        tF, tJ = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
        print("CG Gupta %dx%d SUPG=%s: dofs %.2fM nnz %.2fM setup %.1fs  F %.2f ms  J %.2f ms  -> %.0f MDOF/s (F+J), J writes %.0f GB/s"
              % (nx, 2 * ny, supg, eng.ndof / 1e6, eng.nnz / 1e6, t1 - t0, tF, tJ, eng.ndof / (tF + tJ) / 1e3,
                 8 * eng.nnz / tJ / 1e6), flush=True)
        del eng, s
        torch.cuda.empty_cache()
