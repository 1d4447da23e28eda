"""Full setup_solver() + solve() of the cfg2 MMS case at increasing mesh sizes (run on the GPU box)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import product_cases
from echemfem_b200.fd import errornorm

sizes = [int(a) for a in sys.argv[1:]] or [64, 256]
PC = os.environ.get("PC", "ilu")
CUSTOM = {}
if os.environ.get("COARSE"):
    CUSTOM["pc_coarse_dofs"] = int(os.environ["COARSE"])
if os.environ.get("COARSE_TYPE"):
    CUSTOM["pc_coarse_type"] = os.environ["COARSE_TYPE"]
if os.environ.get("KSP_RTOL"):
    CUSTOM["ksp_rtol"] = float(os.environ["KSP_RTOL"])
for N in sizes:
    t0 = time.time()
    mesh = None
    if os.environ.get("TILE"):
        from echemfem_b200.mesh import UnitSquareMesh
        mesh = UnitSquareMesh(N, N, tile=int(os.environ["TILE"]))
    s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, D1=0.5e-5, D2=1.0e-5, mesh=mesh)
    s.init_solver_parameters(pc_type=PC)
    s.solver_parameters.update(CUSTOM)
    s.potential_solver_parameters.update(CUSTOM)
    s.solver_parameters.pop("snes_monitor", None)
    t1 = time.time()
    s.setup_solver()
    torch.cuda.synchronize()
    t2 = time.time()
    s.solve()
    torch.cuda.synchronize()
    t3 = time.time()
    e = s._engine
    c1, U = s.u.subfunctions
    print("N=%d dofs=%d build %.1fs presolve %.2fs solve %.2fs  snes_its=%d ksp_its=%s  errC=%.3e errU=%.3e timers=%s" % (
        N, e.ndof, t1 - t0, t2 - t1, t3 - t2, e.last["snes_its"], e.last["ksp_its_per_step"],
        errornorm(s.C1ex, c1), errornorm(s.Uex, U), {k: round(v, 3) for k, v in e.timers.items()}), flush=True)
