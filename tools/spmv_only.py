"""SpMV on the cfg2 matrix, a few launches (target of the ncu capture in profiles/)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import product_cases
from echemfem_b200.newton import NewtonEngine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1408
s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, D1=0.5e-5, D2=1.0e-5)
eng = NewtonEngine(s)
x = s.V.node_coords()
s.u.tensor.copy_(torch.from_numpy(np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3, np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3])).cuda())
eng.jacobian(s.u.tensor)
xv = torch.from_numpy(np.cos(np.arange(eng.ndof) * 0.37)).cuda(); yv = torch.empty_like(xv)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3): eng.spmv(xv, yv)
ev[0].record()
for _ in range(10): eng.spmv(xv, yv)
ev[1].record(); torch.cuda.synchronize()
t = ev[0].elapsed_time(ev[1]) / 10
print("spmv %.3f ms, %.0f GB/s algorithmic (12 nnz + 20 rows)" % (t, (12 * eng.nnz + 20 * eng.ndof) / t / 1e6))
