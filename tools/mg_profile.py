"""Where a multigrid-preconditioned Krylov iteration spends its time at cfg2 size (GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import product_cases
from echemfem_b200.mesh import UnitSquareMesh
from echemfem_b200.newton import NewtonEngine
from echemfem_b200 import capi
from echemfem_b200.capi import ptr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1408
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, D1=0.5e-5, D2=1.0e-5, mesh=UnitSquareMesh(N, N, tile=T))
eng = NewtonEngine(s)
x = s.V.node_coords()
u = np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3, np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3])
s.u.tensor.copy_(torch.from_numpy(u).cuda())
F = eng.residual(s.u.tensor).clone()
eng.jacobian(s.u.tensor)
M = eng.multigrid()
t0 = time.time(); M.setup(); torch.cuda.synchronize(); print("setup %.2f s, levels %d" % (time.time() - t0, M.nlevels))

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

r = F.clone(); z = torch.zeros_like(r)
print("V-cycle total      %.2f ms" % timed(lambda: M.apply(r, z)))
for l, D in enumerate(M.lv[:-1]):
    print("level %d: ndof %9d nblocks %6d  smooth %.3f ms  spmv %.3f ms  restrict %.3f ms  prolong %.3f ms" % (
        l, D.ndof, D.nblocks, timed(lambda: M._smooth(D, D.r if l else r, D.z)), timed(lambda: M._spmv(D, D.z, D.t)),
        timed(lambda: M._restrict(D, M.lv[l + 1], D.t, M.lv[l + 1].r)),
        timed(lambda: M._prolong_add(D, M.lv[l + 1], M.lv[l + 1].z, D.s))))
if M.aux is not None:
    A = M.aux
    st = eng._stream()
    print("aux: restrict + V-cycle %.3f ms   prolong-add %.3f ms   (%d conforming levels, %d vertices, %d nnz)" % (
        timed(lambda: A.restrict_and_solve(M.lv[0].t, st)), timed(lambda: A.prolong_add(z, M.lv[0].s, st)),
        len(A.lv), A.ncg, A.lv[0].vals.numel()))
    t0 = time.time(); A.setup(eng.rowptr, eng.colidx, eng.vals, st); torch.cuda.synchronize()
    print("aux set-up (numeric) %.3f s" % (time.time() - t0))
    print("aux: Q^T A Q kernel %.3f ms" % timed(lambda: capi.check(eng.lib.ech_aux_galerkin(
        A.nv, A.row0.numel(), ptr(A.row0), ptr(A.off), ptr(A.dst), ptr(eng.rowptr), ptr(eng.vals), ptr(A.Wd),
        A.lv[0].vals.numel(), ptr(A.lv[0].vals), st))))
    for l, D in enumerate(A.lv[:-1]):
        print("  aux level %d: n %8d nnz %9d" % (l, D.n, D.vals.numel()))
L = eng._linear_plan(200)
V = L["basis"]; n = eng.ndof; work = L["work"]; hdev = work[3 * n:3 * n + 202]; partial = work[3 * n + 4096:]
for k in (10, 50, 100):
    print("multidot k=%d %.2f ms   multiaxpy %.2f ms" % (k, timed(lambda: capi.check(eng.lib.ech_multidot(n, k, ptr(V), ptr(r), ptr(partial), ptr(hdev), eng._stream()))),
          timed(lambda: capi.check(eng.lib.ech_multiaxpy(n, k, ptr(V), ptr(hdev), -1.0, ptr(z), eng._stream())))))

# does the cycle cost depend on WHICH set-up produced it?  (KSP profile: 13.4 / 18.8 ms per application in
# consecutive Newton iterations)
for rep in range(4):
    eng.jacobian(s.u.tensor)
    M.setup(); torch.cuda.synchronize()
    D = M.lv[0]
    line = "set-up %d: V-cycle %.2f ms  fine smooth %.3f  fine spmv %.3f  level-1 smooth %.3f" % (
        rep, timed(lambda: M.apply(r, z)), timed(lambda: M._smooth(D, r, D.z)), timed(lambda: M._spmv(D, D.z, D.t)),
        timed(lambda: M._smooth(M.lv[1], M.lv[1].r, M.lv[1].z)))
    if M.aux is not None:
        line += "  aux cycle %.3f" % timed(lambda: M.aux.restrict_and_solve(D.t, eng._stream()))
    print(line + "  eager V-cycle %.2f" % timed(lambda: M._apply_eager(r, z)))

# where a set-up (PCSetUp, once per Newton iteration) goes
lib, st = eng.lib, eng._stream()
def factor(D):
    capi.check(lib.ech_bjacobi_ilu0_factor(D.ndof, M.nf, D.N, M.h.n, D.nblocks, ptr(D.block_ptr), ptr(D.rowptr),
                                           ptr(D.colidx), ptr(D.vals), ptr(D.lu), ptr(D.diag), st))
print("set-up parts: " + "  ".join(
    ["factor L%d %.1f ms" % (l, timed(lambda: factor(D), 2)) for l, D in enumerate(M.lv[:-1])]))
flag = torch.zeros(1, dtype=torch.int32, device="cuda")
print("set-up parts (factorisation on local columns): " + "  ".join(
    ["factor L%d %.1f ms" % (l, timed(lambda: D.ilu_plan.factor_async(D.rowptr, D.colidx, D.vals, D.lu, D.diag, ptr(flag), st), 2))
     for l, D in enumerate(M.lv[:-1]) if D.ilu_plan is not None]))
print("set-up parts: " + "  ".join(
    ["pack L%d %.2f ms" % (l, timed(lambda: D.ilu_plan.refresh(D.lu, st), 2)) for l, D in enumerate(M.lv[:-1]) if D.ilu_plan is not None]))
if M.own_galerkin:
    print("set-up parts: " + "  ".join(
        ["galerkin L%d->L%d %.2f ms" % (l, l + 1, timed(lambda: M._galerkin(M.lv[l], M.lv[l + 1]), 2)) for l in range(M.nlevels - 1)]))
C_ = M.lv[-1]
def inv():
    dense = torch.sparse_csr_tensor(C_.rowptr, C_.colidx.to(torch.int64), C_.vals, size=(C_.ndof, C_.ndof)).to_dense()
    return torch.linalg.inv(dense)
print("set-up parts: dense inverse of the coarsest level (%d) %.1f ms" % (C_.ndof, timed(inv, 2)))
t0 = time.time(); M.setup(); torch.cuda.synchronize(); print("whole set-up %.3f s" % (time.time() - t0))
