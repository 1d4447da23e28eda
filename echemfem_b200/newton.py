"""Device Newton solver behind `EchemSolver.setup_solver()` / `solve()`.

Takes the place of `NonlinearVariationalProblem` + `NonlinearVariationalSolver`
(echemfem/solver.py:714-725) and of the PETSc objects underneath: SNES newtonls
with the l2 line search, KSP (F)GMRES, PC bjacobi + sub ILU(0) (solver.py:
938-957, 1179-1189).  All arrays live in HBM; per Newton step the host reads
back only norms and the Hessenberg column of each Krylov iteration.

Deviation, stated: `pc_type="lu"` (MUMPS in the reference, solver.py:1173-1178)
has no device counterpart here; it is served by the same GMRES + block-Jacobi /
ILU(0) path run to `ksp_rtol` 1e-12, which reproduces the Newton iterates to
well below the 1e-8 solution tolerance of BASELINE.json.
"""
import ctypes as C
import math
import os
import sys
import time

import numpy as np
import torch

from . import build, capi
from .capi import ptr
from .compiler import CompiledForm


class ConvergenceError(RuntimeError):
    pass


class _Snes:
    """The few SNES queries the reference makes (solver.py:719-720, 750-758)."""

    def __init__(self, engine):
        self._e = engine
        self.ksp = self

    def setConvergenceHistory(self):
        pass

    def getIterationNumber(self):
        return self._e.last["snes_its"]

    def getLinearSolveIterations(self):
        return self._e.last["ksp_its"]

    def getConvergenceHistory(self):
        return np.array(self._e.last["history"]), np.array(self._e.last["ksp_its_per_step"])


class NewtonEngine:
    def __init__(self, solver, nblocks=None, max_basis_bytes=None):
        if not torch.cuda.is_available():
            raise RuntimeError("echemfem_b200 needs a CUDA device: there is no CPU path")
        self.solver = solver
        self.lib = capi.lib()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        mesh = solver.mesh
        W = solver.W
        self.nf = W.num_sub_spaces()
        self.naux = len(solver._aux)
        V = W.scalar
        self.family = V.family
        self.nloc = V.n
        self.N = V.num_nodes                                   # per-field stride: owned + ghost cells
        self.ndof = self.nf * self.N
        self.ncell_total = mesh.num_cells
        self.ncell = getattr(mesh, "num_owned", mesh.num_cells)
        self.halo = None
        t0 = time.time()
        self.form = CompiledForm(solver.Form, self.nf, self.naux, mesh.dim, V.degree, V.family)
        self.module = build.build_module(self.form.header, self.form.key)
        self.t_compile = time.time() - t0
        self._upload_mesh(mesh)
        self.cg = self._cg_plan(V) if self.family == "CG" else None
        self._create()
        self._pattern()
        self.bc_mask, self.bc_vals = self._dirichlet(V) if self.family == "CG" else (None, None)
        plan = getattr(mesh, "halo_plan", None)
        if plan is not None and plan.size > 1:
            from .partition import HaloExchange
            self.halo = HaloExchange(plan, self.nf, self.nloc, self.ncell_total, self.dev)
        self.snes = _Snes(self)
        self.last = {"snes_its": 0, "ksp_its": 0, "history": [], "ksp_its_per_step": []}
        self.timers = {k: 0.0 for k in ("SNESSolve", "KSPSolve", "PCSetUp", "PCApply", "JacobianEval", "FunctionEval")}
        self.nblocks = nblocks
        self.max_basis_bytes = max_basis_bytes
        self._lin = None

    # -- plans ---------------------------------------------------------------------
    def _upload_mesh(self, mesh):
        dev = self.dev
        nbr, code, slot, slot_self, ncol = mesh.connectivity(self.form.marker_class)
        f = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)
        nc = self.ncell_total
        self.m = {
            "xcell": f(mesh.cell_coords(), torch.float64), "nbr": f(nbr, torch.int32), "code": f(code, torch.uint8),
            "slot": f(slot, torch.uint8), "slot_self": f(slot_self, torch.uint8), "ncol": f(ncol, torch.uint8),
            "vol": torch.empty(nc, dtype=torch.float64, device=dev),
            "diam": torch.empty(nc, dtype=torch.float64, device=dev),
            "area": torch.empty(nc * 2 * mesh.dim, dtype=torch.float64, device=dev)}
        capi.check(self.lib.ech_cell_geometry(mesh.dim, nc, ptr(self.m["xcell"]), ptr(self.m["vol"]),
                                              ptr(self.m["area"]), ptr(self.m["diam"]), self._stream()))

    # -- continuous spaces: node-owner gather plan --------------------------------------
    def _cg_plan(self, V):
        """cell->node map, its transpose, the node adjacency (row pattern) and column slots (host, numpy)."""
        import scipy.sparse as sp
        cn = V.cell_nodes.astype(np.int64)
        nc, n = cn.shape
        Nn = V.num_nodes
        flat = cn.reshape(-1)
        order = np.argsort(flat, kind="stable")
        inc_cell = (order // n).astype(np.int32)
        inc_loc = (order % n).astype(np.uint8)
        inc_ptr = np.concatenate([[0], np.cumsum(np.bincount(flat, minlength=Nn))]).astype(np.int64)
        rows = np.repeat(cn, n, axis=1).reshape(-1)
        cols = np.tile(cn, (1, n)).reshape(-1)
        Aadj = sp.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(Nn, Nn))
        Aadj.sum_duplicates()
        Aadj.sort_indices()
        adj_ptr = Aadj.indptr.astype(np.int64)
        adj = Aadj.indices.astype(np.int64)
        adj_count = np.diff(adj_ptr).astype(np.int32)
        keys = np.repeat(np.arange(Nn, dtype=np.int64), adj_count) * Nn + adj
        inc_node = flat[order]
        want = inc_node[:, None] * Nn + cn[inc_cell]                          # (ninc, n)
        pos = np.searchsorted(keys, want.reshape(-1)).reshape(want.shape) - adj_ptr[inc_node][:, None]
        f = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev, dtype=dt)
        return {"Nn": Nn, "adj_ptr": adj_ptr, "adj": adj, "adj_count_h": adj_count,
                "cell_nodes": f(cn, torch.int32), "inc_ptr": f(inc_ptr, torch.int64), "inc_cell": f(inc_cell, torch.int32),
                "inc_loc": f(inc_loc, torch.uint8), "colslot": f(pos.astype(np.uint8), torch.uint8),
                "adj_count": f(adj_count, torch.int32)}

    def _cg_pattern(self):
        """Field-blocked AIJ pattern of the CG space: row (i, r) couples, for every field j
        the Jacobian form couples to, the nodes of all cells around node r, ascending."""
        cg = self.cg
        Nn, adj_ptr, adj, cnt = cg["Nn"], cg["adj_ptr"], cg["adj"], cg["adj_count_h"].astype(np.int64)
        own = self.form.own_mask
        lens = np.concatenate([own[i].sum() * cnt for i in range(self.nf)])
        rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        colidx = np.empty(int(rowptr[-1]), dtype=np.int32)
        within = np.arange(len(adj), dtype=np.int64) - np.repeat(adj_ptr[:-1], cnt)
        for i in range(self.nf):
            start = np.repeat(rowptr[i * Nn:(i + 1) * Nn], cnt)
            seg = 0
            for j in range(self.nf):
                if own[i, j]:
                    colidx[start + seg * np.repeat(cnt, cnt) + within] = (j * Nn + adj).astype(np.int32)
                    seg += 1
        return rowptr, colidx

    def _dirichlet(self, V):
        """Strong DirichletBC dofs and values from the solver's bcs list (solver.py:1531, 2009, 2215)."""
        from .function import evaluate_expression, Function
        from . import symbolic as S
        mesh = self.solver.mesh
        d = mesh.dim
        Nn = V.num_nodes
        mask = np.zeros(self.ndof, dtype=np.uint8)
        vals = np.zeros(self.ndof)
        ref = V.ref_nodes()
        coords = V.node_coords()
        for kind, field, value, ids in self.solver.bcs:
            if kind != "dirichlet" or ids is None:
                continue
            ids = (ids,) if isinstance(ids, int) else tuple(ids)
            on = (mesh.nbr_cell < 0) & np.isin(mesh.ext_marker, ids)
            cs, lfs = np.nonzero(on)
            for lf in np.unique(lfs):
                k, sde = divmod(int(lf), 2)
                loc = np.nonzero(np.abs(ref[:, k] - sde) < 1e-14)[0]
                nodes = np.unique(V.cell_nodes[cs[lfs == lf]][:, loc].reshape(-1))
                if isinstance(value, Function):
                    v = value.numpy()[nodes]
                elif isinstance(value, (int, float)):
                    v = np.full(len(nodes), float(value))
                else:
                    v = evaluate_expression(value, coords[nodes])
                mask[field * Nn + nodes] = 1
                vals[field * Nn + nodes] = v
        if not mask.any():
            return None, None
        return torch.from_numpy(mask).to(self.dev), torch.from_numpy(vals).to(self.dev)

    def _create(self):
        mesh = self.solver.mesh
        m = self.m
        cg = self.cg
        z = None
        self.desc = capi.MeshDesc(mesh.dim, self.solver.W.scalar.degree, self.nf, self.naux,
                                  self.ncell, self.ncell_total,
                                  ptr(m["xcell"]), ptr(m["nbr"]), ptr(m["code"]), ptr(m["slot"]),
                                  ptr(m["slot_self"]), ptr(m["ncol"]), ptr(m["vol"]), ptr(m["diam"]), ptr(m["area"]),
                                  cg["Nn"] if cg else 0, ptr(cg["cell_nodes"]) if cg else z,
                                  ptr(cg["inc_ptr"]) if cg else z, ptr(cg["inc_cell"]) if cg else z,
                                  ptr(cg["inc_loc"]) if cg else z, ptr(cg["colslot"]) if cg else z,
                                  ptr(cg["adj_count"]) if cg else z)
        h = C.c_void_p()
        capi.check(self.lib.ech_create(self.module.encode(), C.byref(self.desc), C.byref(h)))
        self.h = h

    def set_variant(self, v):
        """0: cooperative kernels (default); 1: row-per-thread kernels (cross-check / large elements)."""
        mod = C.CDLL(self.module)
        mod.ech_mod_set_variant.argtypes = [C.c_int]
        mod.ech_mod_set_variant.restype = None
        mod.ech_mod_set_variant(int(v))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _pattern(self):
        n = C.c_int64()
        capi.check(self.lib.ech_num_rows(self.h, C.byref(n)), self.h)
        assert n.value == self.ndof
        if self.cg is not None:
            rp, ci = self._cg_pattern()
            self.rowptr = torch.from_numpy(rp).to(self.dev)
            self.colidx = torch.from_numpy(ci).to(self.dev)
            self.nnz = int(rp[-1])
        else:
            self.rowptr = torch.empty(self.ndof + 1, dtype=torch.int64, device=self.dev)
            nnz = C.c_int64()
            capi.check(self.lib.ech_pattern_rowptr(self.h, ptr(self.rowptr), C.byref(nnz), self._stream()), self.h)
            self.nnz = nnz.value
            self.colidx = torch.empty(self.nnz, dtype=torch.int32, device=self.dev)
            capi.check(self.lib.ech_pattern_colidx(self.h, ptr(self.rowptr), ptr(self.colidx), self._stream()), self.h)
        self.vals = torch.empty(self.nnz, dtype=torch.float64, device=self.dev)
        self.F = torch.zeros(self.ndof, dtype=torch.float64, device=self.dev)

    # -- device state -----------------------------------------------------------------
    def _aux_tensor(self):
        if self.naux == 0:
            return None
        return torch.cat([fn.tensor for fn in self.solver._aux]).contiguous()

    def _consts(self):
        return torch.from_numpy(self.form.constant_values()).to(self.dev)

    # -- SNES callbacks -----------------------------------------------------------------
    def residual(self, u, out=None, aux=None, consts=None):
        out = self.F if out is None else out
        aux = self._aux_tensor() if aux is None else aux
        consts = self._consts() if consts is None else consts
        if self.halo is not None:
            self.halo.exchange(u)
        capi.check(self.lib.ech_residual(self.h, ptr(u), ptr(aux), ptr(consts), ptr(out), self._stream()), self.h)
        return out

    def jacobian(self, u, aux=None, consts=None):
        aux = self._aux_tensor() if aux is None else aux
        consts = self._consts() if consts is None else consts
        if self.halo is not None:
            self.halo.exchange(u)
        capi.check(self.lib.ech_jacobian(self.h, ptr(u), ptr(aux), ptr(consts), ptr(self.rowptr), ptr(self.vals),
                                         self._stream()), self.h)
        return self.vals

    def residual_jacobian(self, u, out=None, aux=None, consts=None):
        """F(u) and J(u) in one pass (`ech_residual_jacobian`)."""
        out = self.F if out is None else out
        aux = self._aux_tensor() if aux is None else aux
        consts = self._consts() if consts is None else consts
        if self.halo is not None:
            self.halo.exchange(u)
        capi.check(self.lib.ech_residual_jacobian(self.h, ptr(u), ptr(aux), ptr(consts), ptr(self.rowptr),
                                                  ptr(self.vals), ptr(out), self._stream()), self.h)
        return out, self.vals

    def spmv(self, x, y):
        capi.check(self.lib.ech_spmv(self.ndof, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), ptr(x), ptr(y),
                                     self._stream()))
        return y

    def csr_host(self):
        """(rowptr, colidx, vals) copied to the host -- for tests and small problems only."""
        return self.rowptr.cpu().numpy(), self.colidx.cpu().numpy(), self.vals.cpu().numpy()

    # -- linear solver plan ----------------------------------------------------------------
    def _linear_plan(self, restart):
        if self._lin is not None and self._lin["restart"] >= restart:
            return self._lin
        # preconditioner blocks are ranges of "units": cells (all dofs of the cell) for DG, nodes for CG
        nunits = self.ncell if self.cg is None else self.N
        per = 64 if self.cg is None else 256
        nblocks = self.nblocks
        hint = getattr(self.solver.mesh, "block_hint", None)
        if nblocks is None and hint and self.cg is None:
            nblocks = max(1, nunits // int(hint))          # tile-numbered mesh: one block per tile
        if nblocks is None:
            # one warp factors / solves one block: enough blocks to fill the machine, >= ~256 rows per field each
            nblocks = int(max(1, min(nunits // per, 148 * 32)))
        edges = np.linspace(0, nunits, nblocks + 1).astype(np.int64)
        self._unit = self.nloc if self.cg is None else 1
        free, _ = torch.cuda.mem_get_info()
        cap = self.max_basis_bytes if self.max_basis_bytes is not None else int(0.5 * free)
        m = int(max(2, min(restart, cap // (8 * self.ndof) - 1)))
        if m < restart:
            print("*** WARNING: ksp_gmres_restart %d capped to %d by device memory" % (restart, m))
        dev = self.dev
        from .ilu import IluPlan
        block_ptr = torch.from_numpy(edges).to(dev)
        plan = None
        if os.environ.get("ECH_ILU_PLANNED", "1") != "0":
            # level-scheduled triangular solves: the plan depends on the pattern and the blocks only
            plan = IluPlan(self.lib, dev, self._stream(), self.ndof, self.nf, self.N, self._unit, edges, block_ptr,
                           self.rowptr, self.colidx, self.nnz)
            if not plan.ok:
                plan = None
        self._lin = {
            "restart": m, "nblocks": nblocks, "ilu_plan": plan,
            "block_ptr": block_ptr,
            "lu": torch.empty(self.nnz, dtype=torch.float64, device=dev),
            "diag": torch.empty(self.ndof, dtype=torch.int32, device=dev),
            "basis": torch.empty((m + 1) * self.ndof, dtype=torch.float64, device=dev),
            "work": torch.empty(3 * self.ndof + 4096 + 2 * (m + 2) * 1024, dtype=torch.float64, device=dev),
            "dx": torch.zeros(self.ndof, dtype=torch.float64, device=dev),
        }
        return self._lin

    def _coarse_plan(self, target):
        """Aggregation map of the optional coarse level: dofs are binned by their coordinates into boxes, one
        coarse dof per (field, non-empty box), piecewise-constant prolongation (the nodal bases are partitions of
        unity).  Not an option set of the reference: it stands in for the coarse levels of its gmg / amg sets
        (solver.py:1190-1213) where block-Jacobi/ILU(0) alone needs O(N) Krylov iterations."""
        C_ = getattr(self, "_coarse", None)
        if C_ is not None and C_["target"] == target:
            return C_
        from .coarse import box_aggregates
        agg, nc = box_aggregates(self.solver.W.scalar.node_coords(), self.nf, target)
        dev = self.dev
        self._coarse = {
            "target": target, "nc": nc, "agg": torch.from_numpy(agg).to(dev),
            "Ac": torch.empty((nc, nc), dtype=torch.float64, device=dev),
            "inv": None,
            "work": torch.empty(2 * nc + self.ndof, dtype=torch.float64, device=dev),
        }
        return self._coarse

    def _global_coarse_plan(self, target):
        """Coarse space shared by all ranks (multi-GPU): the same box aggregation on the GLOBAL bounding box.
        Block-Jacobi across ranks (one block per rank, as PETSc's bjacobi) cannot move the smooth global modes of
        the potential equation; the Galerkin operator of this space is summed over the ranks (one all-reduce per
        set-up), inverted redundantly, and applied after the rank-local preconditioner (one all-reduce of nc
        numbers per Krylov iteration).  Ghost dofs take part in the Galerkin product (columns) only."""
        import torch.distributed as dist
        from .coarse import box_aggregates_global
        C_ = getattr(self, "_gcoarse", None)
        if C_ is not None and C_["target"] == target:
            return C_
        dev = self.dev
        x = self.solver.W.scalar.node_coords()
        d = x.shape[1]
        ext = torch.tensor(np.concatenate([x.min(axis=0), -x.max(axis=0)]), dtype=torch.float64, device=dev)
        dist.all_reduce(ext, op=dist.ReduceOp.MIN)
        ext = ext.cpu().numpy()
        agg, nc, _ = box_aggregates_global(x, self.nf, target, ext[:d], -ext[d:])
        no = self.ncell * self.nloc
        owned = np.zeros(self.ndof, dtype=bool)
        for f in range(self.nf):
            owned[f * self.N:f * self.N + no] = True
        present = torch.zeros(nc, dtype=torch.float64, device=dev)
        present[torch.from_numpy(np.unique(agg[owned]).astype(np.int64)).to(dev)] = 1.0
        dist.all_reduce(present)
        keep = (present.cpu().numpy() > 0)
        renum = np.cumsum(keep) - 1                      # globally empty boxes get no coarse dof
        ncc = int(keep.sum())
        agg = renum[agg].astype(np.int32)
        agg_owned = np.where(owned, agg, ncc).astype(np.int32)     # ghost dofs -> bin `ncc`, never read back
        self._gcoarse = {
            "target": target, "nc": ncc, "agg": torch.from_numpy(agg).to(dev),
            "agg_owned": torch.from_numpy(agg_owned).to(dev),
            "Ac": torch.empty((ncc, ncc), dtype=torch.float64, device=dev), "inv": None,
            "rc": torch.zeros(ncc + 1, dtype=torch.float64, device=dev),
            "yc": torch.zeros(ncc + 1, dtype=torch.float64, device=dev),
            "res": torch.empty(self.ndof, dtype=torch.float64, device=dev),
        }
        return self._gcoarse

    def _global_coarse_setup(self, target):
        import torch.distributed as dist
        Cz = self._global_coarse_plan(target)
        capi.check(self.lib.ech_coarse_galerkin(self.ndof, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals),
                                                ptr(Cz["agg"]), Cz["nc"], ptr(Cz["Ac"]), self._stream()))
        dist.all_reduce(Cz["Ac"])
        # dense coarse inverse: a library call (cuSOLVER through torch), once per Newton iteration, on every rank
        Cz["inv"] = torch.linalg.inv(Cz["Ac"]).contiguous()
        return Cz

    def _global_coarse_correct(self, Cz, r, z):
        """z += P Ac^-1 P^T (r - A z)   (multiplicative correction after the rank-local preconditioner)."""
        import torch.distributed as dist
        lib, st, n = self.lib, self._stream(), self.ndof
        self.halo.exchange(z)
        capi.check(lib.ech_spmv(n, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), ptr(z), ptr(Cz["res"]), st))
        capi.check(lib.ech_axpby(n, 1.0, ptr(r), -1.0, ptr(Cz["res"]), st))             # res = r - A z
        Cz["rc"].zero_()
        capi.check(lib.ech_coarse_restrict(n, ptr(Cz["agg_owned"]), Cz["nc"] + 1, ptr(Cz["res"]), ptr(Cz["rc"]), st))
        dist.all_reduce(Cz["rc"][:Cz["nc"]])
        capi.check(lib.ech_dense_gemv(Cz["nc"], Cz["nc"], ptr(Cz["inv"]), ptr(Cz["rc"]), ptr(Cz["yc"]), st))
        capi.check(lib.ech_coarse_prolong_add(n, ptr(Cz["agg_owned"]), ptr(Cz["yc"]), ptr(z), st))
        z.mul_(self._owned_mask())

    def _global_conforming(self, M):
        """The conforming coarse space shared by all ranks (built once), or None if the mesh has no global axes
        or the problem no potential field."""
        if not hasattr(self, "_gconf"):
            self._gconf = None
            axes = getattr(self.solver.mesh, "global_axes", None)
            if axes is not None and getattr(M, "aux", None) is not None:
                from .auxspace import GlobalConformingCoarse
                try:
                    self._gconf = GlobalConformingCoarse(self, M.aux.fields, axes)
                except NotImplementedError:
                    self._gconf = None
        return self._gconf

    def multigrid(self, **kw):
        """The geometric multigrid hierarchy of this mesh (built once), or None if the mesh has none."""
        if not hasattr(self, "_mg"):
            self._mg = None
            try:
                from .multigrid import Multigrid
                # validated for Q1; higher degrees need a p-coarsening level first (the reference's PMGPC,
                # examples/mms_scaling.py:211-238), until then they keep the one-level / aggregation path
                local_ok = self.halo is None or hasattr(self.solver.mesh, "owned_axes")
                if local_ok and self.cg is None and self.solver.W.scalar.degree == 1:
                    self._mg = Multigrid(self, **kw)
            except NotImplementedError:
                self._mg = None
        return self._mg

    def linear_solve(self, rhs, rtol, atol, restart, max_it, monitor=False, coarse=0, coarse_mode=3, mg=False,
                     refine=None):
        """Solve J dx = rhs with right-preconditioned GMRES; returns (dx, iterations, converged).
        refine: Gram-Schmidt refinement, 0 never / 1 if needed / 2 always; None: by tolerance -- classical
        Gram-Schmidt without refinement (PETSc's default) loses orthogonality like cond^2 * eps, which the
        reference's iterative option sets (ksp_rtol 1e-2 ... 1e-10 on preconditioned systems) tolerate and the
        1e-12 solves that stand in for its direct solver do not."""
        if refine is None:
            refine = 1 if rtol < 1e-8 else 0
        L = self._linear_plan(restart)
        restart = int(max(1, min(restart, L["restart"])))        # a cached plan may hold a longer basis
        M = self.multigrid() if mg else None
        if M is not None:
            t0 = time.time()
            M.setup()
            torch.cuda.synchronize()
            self.timers["PCSetUp"] += time.time() - t0
            pre = M.apply
            if self.halo is not None and coarse:
                # coarse space shared by all ranks, applied after the rank-local V-cycles: continuous Q1 on the
                # global vertex grid for the potential fields (auxspace.GlobalConformingCoarse) where the mesh
                # knows its global axes, else (or with ECH_GCOARSE=box / both) the box aggregation of section 3.4
                mode = os.environ.get("ECH_GCOARSE", "cg")
                G = self._global_conforming(M) if mode in ("cg", "both") else None
                t0 = time.time()
                Cz = self._global_coarse_setup(int(coarse)) if (G is None or mode == "both") else None
                if G is not None:
                    G.setup(self._stream())
                torch.cuda.synchronize()
                self.timers["PCSetUp"] += time.time() - t0

                def pre(src, dst):
                    M.apply(src, dst)
                    if Cz is not None:
                        self._global_coarse_correct(Cz, src, dst)
                    if G is not None:
                        G.correct(src, dst)
                        dst.mul_(self._owned_mask())
            t0 = time.time()
            out = self._gmres(rhs, rtol, atol, max_it, monitor, precond=pre, restart=restart, refine=refine)
            torch.cuda.synchronize()
            self.timers["KSPSolve"] += time.time() - t0
            return out
        t0 = time.time()
        if L["ilu_plan"] is not None and os.environ.get("ECH_ILU_FACTOR_PLANNED", "1") != "0":
            L["ilu_plan"].factor(self.rowptr, self.colidx, self.vals, L["lu"], L["diag"], self._stream())
        else:
            capi.check(self.lib.ech_bjacobi_ilu0_factor(self.ndof, self.nf, self.N, self._unit, L["nblocks"],
                                                        ptr(L["block_ptr"]), ptr(self.rowptr), ptr(self.colidx),
                                                        ptr(self.vals), ptr(L["lu"]), ptr(L["diag"]), self._stream()))
            if L["ilu_plan"] is not None:
                L["ilu_plan"].refresh(L["lu"], self._stream())
        Cz = None
        if coarse and self.halo is None:
            Cz = self._coarse_plan(int(coarse))
            capi.check(self.lib.ech_coarse_galerkin(self.ndof, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals),
                                                    ptr(Cz["agg"]), Cz["nc"], ptr(Cz["Ac"]), self._stream()))
            # the dense coarse solve is a library call (cuSOLVER through torch), once per Newton iteration
            Cz["inv"] = torch.linalg.inv(Cz["Ac"]).contiguous()
        torch.cuda.synchronize()
        self.timers["PCSetUp"] += time.time() - t0
        if self.halo is not None:
            t0 = time.time()
            out = self._gmres(rhs, rtol, atol, max_it, monitor, restart=restart, refine=refine)
            torch.cuda.synchronize()
            self.timers["KSPSolve"] += time.time() - t0
            return out
        opts = capi.GmresOpts(rtol, atol, restart, max_it, 1, 1 if monitor else 0, self.nf, self._unit,
                              self.N, L["nblocks"], ptr(L["block_ptr"]), ptr(L["lu"]), ptr(L["diag"]),
                              None, 0, 0, None, None, None, None, None, 0, 0, int(refine), 0)
        if L["ilu_plan"] is not None:
            L["ilu_plan"].set_opts(opts)
        if Cz is not None:
            opts.precond = int(coarse_mode)
            opts.agg, opts.nc = ptr(Cz["agg"]), Cz["nc"]
            opts.coarse_inv, opts.coarse_work = ptr(Cz["inv"]), ptr(Cz["work"])
        L["dx"].zero_()
        its = C.c_int32()
        rn = C.c_double()
        t0 = time.time()
        rc = self.lib.ech_gmres(self.ndof, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), ptr(rhs), ptr(L["dx"]),
                                C.byref(opts), ptr(L["basis"]), ptr(L["work"]), C.byref(its), C.byref(rn),
                                self._stream())
        torch.cuda.synchronize()
        self.timers["KSPSolve"] += time.time() - t0
        if rc not in (0, -5):
            capi.check(rc)
        return L["dx"], its.value, rc == 0

    def _gmres(self, b, rtol, atol, max_it, monitor=False, precond=None, restart=None, refine=0):
        """Right-preconditioned restarted GMRES driven from Python: the same algorithm as `ech_gmres`, with
        (across ranks) the ghost-dof exchange in front of every SpMV and the Gram-Schmidt dot products
        all-reduced over NCCL (PETSc: VecScatter + MPI_Allreduce inside KSPSolve), and with a pluggable
        preconditioner (multigrid V-cycle).  Preconditioner blocks never cross ranks.

        ONE reduction (one all-reduce, one device -> host read-back) per iteration: the fused multi-dot returns
        h = V^T w together with w.w, and |w - V h|^2 = w.w - |h|^2 (classical Gram-Schmidt without refinement is
        PETSc's default, KSP_GMRES_CGS_REFINE_NEVER); a refinement pass runs only where that difference cancels
        (or per `refine`: 1 if needed, 2 always).  Convergence claimed by the recurrence is confirmed with the true
        residual."""
        import torch.distributed as dist
        distributed = self.halo is not None
        L = self._lin
        n = self.ndof
        m = L["restart"] if restart is None else int(max(1, min(restart, L["restart"])))
        lib, st = self.lib, self._stream()
        V = L["basis"]
        work = L["work"]
        w, zt, t2 = work[:n], work[n:2 * n], work[2 * n:3 * n]
        hdev = work[3 * n:3 * n + m + 2]
        partial = work[3 * n + 4096:]
        x = L["dx"]
        x.zero_()
        own = self._owned_mask() if distributed else None
        self.ksp_reductions = 0

        def mdot(k, vec):
            # ghost entries of Krylov vectors are zero (rows of ghost cells are empty), so plain sums are owned sums
            capi.check(lib.ech_multidot(n, k, ptr(V), ptr(vec), ptr(partial), ptr(hdev), st))
            if distributed:
                dist.all_reduce(hdev[:k + 1])
            self.ksp_reductions += 1
            return hdev[:k + 1].cpu().numpy()

        if precond is None:
            def precond(src, dst):
                dst.zero_()
                if L["ilu_plan"] is not None:
                    L["ilu_plan"].apply(L["lu"], src, dst, st)
                    return
                capi.check(lib.ech_bjacobi_ilu0_apply(n, self.nf, self.N, self._unit, L["nblocks"],
                                                      ptr(L["block_ptr"]), ptr(self.rowptr), ptr(self.colidx),
                                                      ptr(L["lu"]), ptr(L["diag"]), ptr(src), ptr(dst), st))

        def matvec(src, dst):
            if distributed:
                self.halo.exchange(src)
            capi.check(lib.ech_spmv(n, ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), ptr(src), ptr(dst), st))

        def owned_sq(vec):
            if not distributed:
                return max(mdot(0, vec)[0], 0.0)
            torch.mul(vec, own, out=t2)
            capi.check(lib.ech_multidot(n, 0, ptr(V), ptr(t2), ptr(partial), ptr(hdev), st))
            dist.all_reduce(hdev[:1])
            self.ksp_reductions += 1
            return max(float(hdev[0].item()), 0.0)

        # refinement threshold of "if needed": a second Gram-Schmidt pass when |w - V h|^2 < eta2 |w|^2
        # (0.5 = PETSc's 1/sqrt(2) criterion on the norms)
        eta2 = float(os.environ.get("ECH_GMRES_REFINE_ETA2", "0.5"))
        # ech_multiaxpy_dot (update of pass one fused with the projection of pass two: three reads of the basis
        # instead of four).  Measured at cfg2 size: SLOWER, 5.6 instead of 3.6 ms per iteration -- one accumulator
        # per basis vector and thread costs the occupancy the streaming kernels live on.  Opt-in.
        fused = os.environ.get("ECH_GMRES_FUSED_CGS2", "0") != "0"
        # ECH_KSP_PROFILE=1: synchronising wall-clock split of the iteration (diagnostic; slows the solve down)
        prof = {"precond": 0.0, "matvec": 0.0, "its": 0, "t0": time.perf_counter()} if os.environ.get("ECH_KSP_PROFILE") else None
        bnorm = math.sqrt(owned_sq(b))
        tol = max(rtol * bnorm, atol)
        its, rnorm, converged = 0, bnorm, False
        H = np.zeros((m + 1, m))
        cs, sn, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
        rank0 = not distributed or self.halo.plan.rank == 0
        verified = None                      # true residual at the last failed verification
        while its < max_it:
            V0 = V[:n]
            t2.copy_(x)
            matvec(t2, w)
            torch.sub(b, w, out=V0)
            if distributed:
                V0.mul_(own)                 # b may carry ghost entries; basis vectors must not
            beta = math.sqrt(max(mdot(0, V0)[0], 0.0))
            rnorm = beta
            if monitor and rank0:
                print("  %3d KSP Residual norm %.12e" % (its, rnorm))
            if beta <= (1.5 * tol if converged else tol):
                converged = True
                break
            if converged:
                # the recurrence claimed convergence, the true residual is not there: restart from it -- unless the
                # previous verification already stood at this level (attainable accuracy of an ill-conditioned
                # system: rtol 1e-12 asks for more than cond * eps allows), then the solve is as good as it gets
                if verified is not None and beta > 0.5 * verified:
                    break
                verified = beta
            converged = False
            V0.mul_(1.0 / beta)
            g[:] = 0.0
            g[0] = beta
            j = 0
            while j < m and its < max_it:
                Vj, Vn = V[j * n:(j + 1) * n], V[(j + 1) * n:(j + 2) * n]
                if prof is not None:
                    torch.cuda.synchronize()
                    tp0 = time.perf_counter()
                    precond(Vj, zt)
                    torch.cuda.synchronize()
                    tp1 = time.perf_counter()
                    matvec(zt, w)
                    torch.cuda.synchronize()
                    tp2 = time.perf_counter()
                    prof["precond"] += tp1 - tp0
                    prof.setdefault("pc", []).append(1e3 * (tp1 - tp0))
                    prof["matvec"] += tp2 - tp1
                    prof["its"] += 1
                else:
                    precond(Vj, zt)
                    matvec(zt, w)
                h = mdot(j + 1, w)
                H[:j + 1, j] = h[:j + 1]
                ww = h[j + 1]
                hn2 = ww - float(np.dot(h[:j + 1], h[:j + 1]))
                second = refine == 2 or (refine == 1 and hn2 < eta2 * ww) or not (hn2 > 1e-6 * ww)
                if second and fused and j + 1 <= 64:
                    # the update of the first pass and the projection of the second in one sweep over the basis
                    capi.check(lib.ech_multiaxpy_dot(n, j + 1, ptr(V), ptr(hdev), -1.0, ptr(w), ptr(partial),
                                                     ptr(hdev), st))
                    if distributed:
                        dist.all_reduce(hdev[:j + 2])
                    self.ksp_reductions += 1
                    h = hdev[:j + 2].cpu().numpy()
                else:
                    capi.check(lib.ech_multiaxpy(n, j + 1, ptr(V), ptr(hdev), -1.0, ptr(w), st))
                    if second:
                        h = mdot(j + 1, w)
                if second:
                    H[:j + 1, j] += h[:j + 1]
                    capi.check(lib.ech_multiaxpy(n, j + 1, ptr(V), ptr(hdev), -1.0, ptr(w), st))
                    hn2 = h[j + 1] - float(np.dot(h[:j + 1], h[:j + 1]))
                    if not (hn2 > 1e-6 * h[j + 1]):
                        hn2 = mdot(0, w)[0]
                hn = math.sqrt(max(hn2, 0.0))
                H[j + 1, j] = hn
                if hn > 0.0:
                    torch.mul(w, 1.0 / hn, out=Vn)
                for i in range(j):
                    a_, b_ = H[i, j], H[i + 1, j]
                    H[i, j] = cs[i] * a_ + sn[i] * b_
                    H[i + 1, j] = -sn[i] * a_ + cs[i] * b_
                a_, b_ = H[j, j], H[j + 1, j]
                den = math.hypot(a_, b_)
                cs[j], sn[j] = (1.0, 0.0) if den == 0.0 else (a_ / den, b_ / den)
                H[j, j], H[j + 1, j] = den, 0.0
                g[j + 1] = -sn[j] * g[j]
                g[j] = cs[j] * g[j]
                its += 1
                rnorm = abs(g[j + 1])
                if monitor and rank0:
                    print("  %3d KSP Residual norm %.12e" % (its, rnorm))
                j += 1
                if rnorm <= tol or hn == 0.0:
                    converged = rnorm <= tol
                    break
            k = j
            if k == 0:
                break
            y = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
            hdev[:k].copy_(torch.from_numpy(y).to(self.dev))
            t2.zero_()
            capi.check(lib.ech_multiaxpy(n, k, ptr(V), ptr(hdev), 1.0, ptr(t2), st))
            precond(t2, zt)
            x.add_(zt)
            # converged by the recurrence: the next pass of the loop checks the true residual and restarts if needed
        if distributed:
            x.mul_(own)      # ghost entries of the update are refreshed by the next exchange
        if prof is not None and rank0 and prof["its"]:
            torch.cuda.synchronize()
            tot = time.perf_counter() - prof["t0"]
            k = prof["its"]
            print("KSP profile: %d its, %.2f ms/it total; preconditioner %.2f ms, SpMV (+exchange) %.2f ms, "
                  "orthogonalisation + host %.2f ms" % (k, 1e3 * tot / k, 1e3 * prof["precond"] / k, 1e3 * prof["matvec"] / k,
                                                       1e3 * (tot - prof["precond"] - prof["matvec"]) / k), file=sys.stderr)
            pc = prof["pc"]
            print("KSP profile: preconditioner ms per application: first %s ... min %.2f median %.2f max %.2f" % (
                ["%.1f" % v for v in pc[:6]], min(pc), sorted(pc)[len(pc) // 2], max(pc)), file=sys.stderr)
        return x, its, converged

    # -- Newton -----------------------------------------------------------------------------
    def _owned_mask(self):
        if not hasattr(self, "_owned"):
            w = torch.zeros(self.ndof, dtype=torch.float64, device=self.dev)
            no = self.ncell * self.nloc
            for f in range(self.nf):
                w[f * self.N:f * self.N + no] = 1.0
            self._owned = w
        return self._owned

    def _norm(self, v):
        if self.halo is not None:
            # owned entries only, summed over the ranks (PETSc VecNorm on an MPI Vec)
            import torch.distributed as dist
            t = (v * v * self._owned_mask()).sum().reshape(1)
            dist.all_reduce(t)
            return math.sqrt(max(float(t.item()), 0.0))
        r = C.c_double()
        if self._lin is not None:
            work = self._lin["work"]
        else:
            if not hasattr(self, "_dotwork"):
                self._dotwork = torch.empty(8 + 2 * 1024, dtype=torch.float64, device=self.dev)
            work = self._dotwork
        capi.check(self.lib.ech_dot(v.numel(), ptr(v), ptr(v), ptr(work), C.byref(r), self._stream()))
        return math.sqrt(max(r.value, 0.0))

    def solve(self, frozen_fields=None, parameters=None):
        P = dict(parameters or {})
        rtol = P.get("snes_rtol", 1e-8)
        atol = P.get("snes_atol", 1e-50)
        stol = P.get("snes_stol", 1e-8)
        max_it = P.get("snes_max_it", 50)
        dtol = P.get("snes_divergence_tolerance", 1e4)
        monitor = "snes_monitor" in P
        direct = P.get("pc_type", "lu") == "lu" or P.get("ksp_type") == "preonly"
        ksp_rtol = 1e-12 if direct else P.get("ksp_rtol", 1e-5)
        ksp_atol = P.get("ksp_atol", 1e-50)
        restart = P.get("ksp_gmres_restart", 30) if not direct else 200
        ksp_max_it = P.get("ksp_max_it", 10000)
        linesearch = P.get("snes_linesearch_type", "bt")
        # PETSc: -ksp_gmres_cgs_refinement_type refine_never (default) | refine_ifneeded | refine_always
        refine = {None: None, "refine_never": 0, "refine_ifneeded": 1, "refine_always": 2}[
            P.get("ksp_gmres_cgs_refinement_type")]
        # coarse level: on by default where the reference would factor directly ("lu") and the system is large
        coarse = P.get("pc_coarse_dofs", 4096 if (direct and self.ndof >= 100000) else 0)
        coarse_mode = 2 if P.get("pc_coarse_type", "multiplicative") == "additive" else 3
        # multigrid: the reference's "gmg" option set (pc_type mg), and where it would factor directly on a large
        # tensor mesh; falls back to the one-level / aggregation preconditioners when the mesh has no hierarchy
        mg = P.get("pc_type") == "mg" or (direct and self.ndof >= 100000 and P.get("pc_mg", True))

        u = self.solver.u.tensor
        aux = self._aux_tensor()
        consts = self._consts()
        mask = None
        if frozen_fields:
            mask = torch.zeros(self.ndof, dtype=torch.uint8, device=self.dev)
            for f in frozen_fields:
                mask[f * self.N:(f + 1) * self.N] = 1
        if self.bc_mask is not None:
            # strong Dirichlet rows (CG): the state carries the boundary values, the rows become identity;
            # values are re-evaluated because they may depend on Constants (time, applied potential)
            self.bc_mask, self.bc_vals = self._dirichlet(self.solver.W.scalar)
            sel = self.bc_mask.bool()
            if frozen_fields:
                for f in frozen_fields:
                    sel[f * self.N:(f + 1) * self.N] = False
            u[sel] = self.bc_vals[sel]
            mask = self.bc_mask.clone() if mask is None else torch.maximum(mask, self.bc_mask)
        st = self._stream()

        def resid(x, out):
            t0 = time.time()
            self.residual(x, out, aux, consts)
            if mask is not None:
                capi.check(self.lib.ech_constrain_rows(self.ndof, ptr(mask), None, None, None, ptr(out), st))
            torch.cuda.synchronize()
            self.timers["FunctionEval"] += time.time() - t0
            return out

        tstart = time.time()
        F = resid(u, self.F)
        fnorm = self._norm(F)
        f0 = fnorm
        hist = [fnorm]
        ksp_hist = []
        if monitor:
            print("  %d SNES Function norm %.12e" % (0, fnorm))
        reason = None
        its = 0
        W1 = torch.zeros_like(u)
        F1 = torch.zeros_like(u)
        F2 = torch.zeros_like(u)
        if fnorm < atol:
            reason = "CONVERGED_FNORM_ABS"
        while reason is None and its < max_it:
            t0 = time.time()
            self.jacobian(u, aux, consts)
            if mask is not None:
                capi.check(self.lib.ech_constrain_rows(self.ndof, ptr(mask), ptr(self.rowptr), ptr(self.colidx),
                                                       ptr(self.vals), None, st))
            torch.cuda.synchronize()
            self.timers["JacobianEval"] += time.time() - t0
            y, kits, ok = self.linear_solve(F, ksp_rtol, ksp_atol, restart, ksp_max_it, coarse=coarse,
                                            coarse_mode=coarse_mode, mg=mg, refine=refine)
            ksp_hist.append(kits)
            if not ok:
                print("*** WARNING: linear solve did not reach ksp_rtol in %d iterations" % kits)
            ynorm = self._norm(y)
            # x_new = u - lambda * y
            lam = 1.0
            if linesearch == "l2":
                lam, Fnew, fnew = self._l2_linesearch(resid, u, y, fnorm, W1, F1, F2)
            else:
                W1.copy_(u)
                capi.check(self.lib.ech_axpby(self.ndof, -1.0, ptr(y), 1.0, ptr(W1), st))
                Fnew = resid(W1, F1)
                fnew = self._norm(Fnew)
            capi.check(self.lib.ech_axpby(self.ndof, -lam, ptr(y), 1.0, ptr(u), st))
            self.F.copy_(Fnew)
            F = self.F
            fnorm = fnew
            its += 1
            hist.append(fnorm)
            if monitor:
                print("  %d SNES Function norm %.12e" % (its, fnorm))
            xnorm = self._norm(u)
            if not math.isfinite(fnorm):
                raise ConvergenceError("Nonlinear solve produced a non-finite residual")
            if fnorm < atol:
                reason = "CONVERGED_FNORM_ABS"
            elif fnorm <= rtol * f0:
                reason = "CONVERGED_FNORM_RELATIVE"
            elif lam * ynorm < stol * xnorm:
                reason = "CONVERGED_SNORM_RELATIVE"
            elif fnorm > dtol * f0:
                raise ConvergenceError("Nonlinear solve diverged (DIVERGED_DTOL)")
        torch.cuda.synchronize()
        self.timers["SNESSolve"] += time.time() - tstart
        self.last = {"snes_its": its, "ksp_its": int(sum(ksp_hist)), "history": hist, "ksp_its_per_step": ksp_hist,
                     "reason": reason}
        if reason is None:
            raise ConvergenceError("Nonlinear solve failed to converge after %d nonlinear iterations "
                                   "(DIVERGED_MAX_IT)" % its)
        if "snes_converged_reason" in P:
            print("Nonlinear solve converged due to %s iterations %d" % (reason, its))
        return self.last

    def _l2_linesearch(self, resid, u, y, fnorm, W, F1, F2):
        """PETSc SNESLineSearchL2, one iteration: secant step on d||F||^2/dlambda."""
        st = self._stream()
        n = self.ndof

        def at(lam, out):
            W.copy_(u)
            capi.check(self.lib.ech_axpby(n, -lam, ptr(y), 1.0, ptr(W), st))
            Fo = resid(W, out)
            return Fo, self._norm(Fo)
        lam, lam_old = 1.0, 0.0
        _, nm = at(0.5 * (lam + lam_old), F2)
        Fl, n1 = at(lam, F1)
        f_old, f_mid, f_new = fnorm * fnorm, nm * nm, n1 * n1
        dl = lam - lam_old
        d_new = (3.0 * f_new - 4.0 * f_mid + f_old) / dl
        d_old = (-3.0 * f_old + 4.0 * f_mid - f_new) / dl
        d2 = (d_new - d_old) / dl
        if d2 == 0.0 or not math.isfinite(d2):
            return lam, Fl, n1
        lam_new = lam - d_new / d2
        if not math.isfinite(lam_new) or lam_new <= 1e-12:
            lam_new = 0.5 * (lam + lam_old)
        lam_new = min(lam_new, 1e8)
        Fn, nn = at(lam_new, F1)
        return lam_new, Fn, nn

    # -- stats (solver.py:729-792) ---------------------------------------------------------------
    def write_stats(self, path, overwrite):
        """One CSV row per solve with the reference's columns (echemfem/solver.py:729-792): counts are summed over
        the ranks (owned cells / dofs only), timers are the maximum over the ranks, rank 0 writes."""
        import pandas
        s = self.solver
        owned_cells = int(self.ncell)
        owned_dofs = int(self.nf * self.ncell * self.nloc) if self.cg is None else int(self.ndof)
        counts = [owned_cells, owned_dofs]
        timers = [self.timers[k] for k in ("SNESSolve", "KSPSolve", "PCSetUp", "PCApply", "JacobianEval", "FunctionEval")]
        nproc, rank = 1, 0
        if self.halo is not None:
            import torch.distributed as dist
            nproc, rank = dist.get_world_size(), dist.get_rank()
            c = torch.tensor(counts, dtype=torch.float64, device=self.dev)
            t = torch.tensor(timers, dtype=torch.float64, device=self.dev)
            dist.all_reduce(c)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            counts, timers = [int(v) for v in c.cpu()], [float(v) for v in t.cpu()]
        if rank != 0:
            return
        data = {"num_processes": nproc, "num_cells": counts[0], "dimension": 3,          # literal in the reference (solver.py:764)
                "degreeC": s.poly_degree, "degreeU": s.poly_degreeU, "dofs": counts[1],
                "name": "bortels_threeion_extruded_3d_nondim", "snes_its": self.last["snes_its"],
                "ksp_its": self.last["ksp_its"], "SNESSolve": timers[0], "KSPSolve": timers[1], "PCSetUp": timers[2],
                "PCApply": timers[3], "JacobianEval": timers[4], "FunctionEval": timers[5],
                "mesh_name": s.mesh.name, "csolver": getattr(s, "csolver", "")}
        path = os.path.abspath(path)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        mode = "w" if overwrite else "a"
        header = True if overwrite else not os.path.exists(path)
        pandas.DataFrame(data, index=[0]).to_csv(path, index=False, mode=mode, header=header)
