"""One multigrid set-up and two eager V-cycles at cfg2 size (for an ncu launch list: `ECH_MG_GRAPH=0`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from echemfem_b200.workloads import cfg2_solver, cfg2_state
from echemfem_b200.mesh import TensorMesh
from echemfem_b200.newton import NewtonEngine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1408
s = cfg2_solver(N, mesh=TensorMesh([np.linspace(0, 1, N + 1)] * 2, tile=8))
eng = NewtonEngine(s)
s.u.tensor.copy_(torch.from_numpy(cfg2_state(s.V.node_coords())).cuda())
F = eng.residual(s.u.tensor).clone()
eng.jacobian(s.u.tensor)
M = eng.multigrid()
M.use_graph = False
M.setup()
if not os.environ.get("ECH_ONE_SETUP"):
    M.setup()                   # second set-up: buffers and plans exist
z = torch.zeros_like(F)
torch.cuda.synchronize()
print("MARK: cycles start")
for _ in range(2):
    M.apply(F, z)
torch.cuda.synchronize()
print("done")
