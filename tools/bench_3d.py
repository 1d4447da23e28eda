"""DG assembly timing on the configs[3] physics (Bortels three-ion, extruded hexes, DQ1 x DG1) at increasing
refinement (GPU box).  r = 2 of the example's hierarchy: 256 x 64 x 32 = 524 288 hexes, 12.6 M dofs."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import product_cases
from echemfem_b200.mesh import ExtrudedMesh
from echemfem_b200.meshes import BortelsMesh
from echemfem_b200.newton import NewtonEngine

for r in [int(a) for a in sys.argv[1:]] or [1]:
    f = 2 ** r
    # bortels_structuredquad_nondim_coarse1664.geo: 22 / 23 / 22 x 17 nodes -> 64 x 16 cells; refined r times
    base = BortelsMesh(21 * f + 1, 22 * f + 1, 21 * f + 1, 16 * f + 1)
    layers, hz = 8 * f, 6.0
    t0 = time.time()
    s = product_cases.Bortels3DSolver(ExtrudedMesh(base, layers, hz / layers), hz)
    eng = NewtonEngine(s)
    t1 = time.time()
    x = s.V.node_coords()
    pert = 1e-2 * np.sin(3 * x[:, 0] + 5 * x[:, 1] + 0.7 * x[:, 2])
    u = np.concatenate([1.0 + pert, 1.0 - 0.5 * pert, 0.1 + pert])
    s.u.tensor.copy_(torch.from_numpy(u).cuda())
    F = torch.zeros_like(s.u.tensor)
    for _ in range(3):
        eng.residual_jacobian(s.u.tensor, out=F)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 10
    ev[0].record()
    for _ in range(reps):
        eng.residual_jacobian(s.u.tensor, out=F)
    ev[1].record()
    torch.cuda.synchronize()
    t = ev[0].elapsed_time(ev[1]) / reps
    print("3D Bortels r=%d: %d hexes, dofs %.2fM, nnz %.1fM (%.1f GB), setup %.1fs, fused F+J %.2f ms -> %.0f MDOF/s, %.0f GB/s of CSR"
          % (r, s.mesh.num_cells, eng.ndof / 1e6, eng.nnz / 1e6, 8 * eng.nnz / 1e9, t1 - t0, t, eng.ndof / t / 1e3,
             8 * eng.nnz / t / 1e6), flush=True)
    del eng, s
    torch.cuda.empty_cache()
