import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch
import mms_cases, product_cases
from oracle.mesh import unit_cube_mesh
from oracle.assemble import Discretization
from echemfem_b200.mesh import UnitSquareMesh, ExtrudedMesh
from echemfem_b200.newton import NewtonEngine
disc = Discretization(unit_cube_mesh(3, 2, 2), mms_cases.adm_problem(1.0))
s = product_cases.AdvectionDiffusionMigrationSolver(0, 1.0, mesh=ExtrudedMesh(UnitSquareMesh(3, 2), 2, 0.5))
eng = NewtonEngine(s)
x = disc.node_coords
pert = 1e-3*np.sin(7*x[:,0]+3*x[:,1]+5*x[:,2])
u = np.concatenate([mms_cases.C1ex(x)+pert, mms_cases.Uex(x)-pert])
s.u.tensor.copy_(torch.from_numpy(u).cuda())
J = disc.jacobian(u)
for v in (0, 1):
    eng.set_variant(v)
    eng.vals.zero_()
    eng.jacobian(s.u.tensor)
    vals = eng.vals.cpu().numpy()
    err = np.abs(vals - J.data)
    bad = np.nonzero(err > 1e-9*np.abs(J.data).max())[0]
    print("variant", v, "max err", err.max(), "nbad", len(bad), "of", len(vals))
    if len(bad):
        rows = np.searchsorted(J.indptr, bad, side='right')-1
        cols = J.indices[bad]
        N = disc.ndof_scalar; n=8
        info = [(int(r//N), int((r%N)//n), int(r%n), int(c//N), int((c%N)//n), int(c%n)) for r,c in zip(rows[:12], cols[:12])]
        print(info)
        import collections
        print(collections.Counter([(int(r//N), int(c//N), int(((r%N)//n)==((c%N)//n))) for r,c in zip(rows, cols)]))
