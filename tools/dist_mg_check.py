"""Newton solve across ranks with the rank-local multigrid V-cycle as block preconditioner, against the exact
MMS fields.  torchrun --nproc-per-node R tools/dist_mg_check.py N   (global mesh: (R*N) x N cells, one strip each)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from echemfem_b200.workloads import cfg2_solver
from echemfem_b200.partition import strip_partition
from echemfem_b200.mesh import TensorMesh

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
axes = [np.linspace(0.0, float(world), world * N + 1), np.linspace(0.0, 1.0, N + 1)]
if world > 1:
    mesh, plan = strip_partition(axes, rank, world, tile=8)
else:
    mesh = TensorMesh(axes, tile=8)
s = cfg2_solver(mesh=mesh)
s.init_solver_parameters(pc_type="lu")
if os.environ.get("ECH_REFINE"):
    s.solver_parameters["ksp_gmres_cgs_refinement_type"] = os.environ["ECH_REFINE"]
if rank:
    s.solver_parameters.pop("snes_monitor", None)
s.setup_solver()
torch.cuda.synchronize()
t0 = time.perf_counter()
s.solve()
torch.cuda.synchronize()
t = time.perf_counter() - t0
e = s._engine
x = s.V.node_coords()
u = s.u.numpy()
no = e.ncell * e.nloc
err = max(np.abs(u[:no] - (np.cos(x[:no, 0]) + np.sin(x[:no, 1]) + 3)).max(),
          np.abs(u[e.N:e.N + no] - (np.sin(x[:no, 0]) + np.cos(x[:no, 1]) + 3)).max())
tt = torch.tensor([err, t], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print("DIST_MG world=%d N=%d dofs/rank=%d snes_its=%d ksp_its=%s solve_s=%.2f max_nodal_err=%.3e mg=%s timers=%s" % (
        world, N, e.nf * no, e.last["snes_its"], e.last["ksp_its_per_step"], tt[1].item(), tt[0].item(),
        getattr(e, "_mg", None) is not None, {k: round(v, 3) for k, v in e.timers.items()}))
if world > 1:
    dist.destroy_process_group()
