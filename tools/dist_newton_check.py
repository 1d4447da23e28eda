"""Distributed Newton solve (torchrun, one rank per GPU) vs the serial solve of the same problem.

Run as:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_newton_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import product_cases
from echemfem_b200.mesh import UnitSquareMesh
from echemfem_b200.partition import partition_mesh

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
kind = sys.argv[2] if len(sys.argv) > 2 else "range"        # "range": contiguous cell ranges; "graph": graph_partition
mesh = UnitSquareMesh(N, N)
# serial reference on this rank's GPU
ss = product_cases.AdvectionDiffusionMigrationSolver(0, 1.0, mesh=mesh)
ss.solver_parameters.pop("snes_monitor", None)
ss.setup_solver()
ss.solve()
u_serial = ss.u.numpy()
# distributed
parts = None
if kind == "graph":
    from echemfem_b200.partition import graph_partition
    parts = graph_partition(mesh, world)
lmesh, plan = partition_mesh(mesh, rank, world, parts=parts)
sd = product_cases.AdvectionDiffusionMigrationSolver(0, 1.0, mesh=lmesh)
if rank != 0:
    sd.solver_parameters.pop("snes_monitor", None)
sd.setup_solver()
sd.solve()
u = sd.u.numpy()
n = 4
Ng, Nl = mesh.num_cells * n, lmesh.num_cells * n
gdof = (lmesh.global_ids[:, None] * n + np.arange(n)[None, :]).reshape(-1)
no = lmesh.num_owned * n
err = 0.0
for f in range(2):
    a = u[f * Nl:f * Nl + no]
    b = u_serial[f * Ng + gdof[:no]]
    err = max(err, np.linalg.norm(a - b) / np.linalg.norm(b))
t = torch.tensor([err], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    e = sd._engine.last
    print("DIST_NEWTON kind=%s world=%d N=%d snes_its=%d ksp_its=%s max_rel_err=%.3e" % (kind, world, N, e["snes_its"], e["ksp_its_per_step"], t.item()))
    assert t.item() < 1e-8
dist.destroy_process_group()
