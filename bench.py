#!/usr/bin/env python
"""Headline benchmark: residual + Jacobian assembly throughput (MDOF/s) of the DG Nernst-Planck MMS case.

Workload (BASELINE.json configs[1]; SURVEY.md 8d "cfg2"): 2D electroneutral Nernst-Planck MMS,
UnitSquareMesh(N, N) quadrilaterals, DG Q1, fields (C1, Phi), D = (0.5e-5, 1e-5) as in
examples/mms_scaling.py:90-93, N = 1408 -> 15.86 M dofs, 634 M non-zeros.
State: nodal interpolant of the exact fields + 1e-3 sin(7x + 3y).

One "step" = one residual evaluation + one Jacobian evaluation (what SNES does once per Newton
iteration, echemfem/solver.py:725).  Prints ONE JSON line (see the task contract).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "assembly_throughput_residual_plus_jacobian"
UNIT = "MDOF/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms; only samples inside the timed region count."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        q = "timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            deadline = time.time() + 5.0
            while not self.rows and time.time() < deadline:      # wait until the sampler is live
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                parts = [p.strip() for p in r.split(",")]
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except Exception:
                    continue
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return sm, mx, reasons
        inside = [r for r in self.rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.04]
        sm, mx, reasons = parse(inside if inside else self.rows)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(inside)}


def ncu_traffic(deg, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the fused kernel, from the committed
    `ncu --set full` capture of this workload (profiles/traffic.json); None when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get("cfg2_Q%d_N%d_fused" % (deg, n), {}).get("dram_bytes")
    except Exception:
        return None


def algorithmic_bytes(mesh, nf, nloc, nnz):
    ndof = nf * getattr(mesh, "num_owned", mesh.num_cells) * nloc
    nvert = mesh.coords.shape[0]
    d = mesh.dim
    BJ = 8 * nnz + 8 * ndof + 8 * d * nvert + 10 * mesh.num_interior_facets + 5 * mesh.num_exterior_facets
    BF = 16 * ndof + 8 * d * nvert + 10 * mesh.num_interior_facets + 5 * mesh.num_exterior_facets
    BS = 12 * nnz + 20 * ndof + 4
    return BJ, BF, BS


def cpu_baseline(sample_n, reps=1):
    """The oracle (numpy port of the reference algorithm) on this host: residual + Jacobian, one core."""
    import mms_cases
    from oracle.mesh import unit_square_mesh
    from oracle.assemble import Discretization
    import numpy as np
    disc = Discretization(unit_square_mesh(sample_n), mms_cases.adm_problem(1.0, D1=0.5e-5, D2=1.0e-5))
    x = disc.node_coords
    u = np.concatenate([mms_cases.C1ex(x), mms_cases.Uex(x)]) + 1e-3 * np.tile(np.sin(7 * x[:, 0] + 3 * x[:, 1]), 2)
    disc.block_mask()
    disc.pattern()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        disc.residual(u)
        disc.jacobian(u)
        best = min(best, time.perf_counter() - t0)
    return disc.ndof / best / 1e6, best, disc.ndof


def _cpu_worker(sample_n):
    v, t, nd = cpu_baseline(sample_n)
    return t, nd


def cpu_baseline_parallel(sample_n, procs, pool=None):
    """The oracle on `procs` host cores at once, one process per core, each assembling its own sample mesh (the
    path shards by cells with no exchange during assembly, as `mpiexec -n procs` would run the reference).
    Returns (aggregate MDOF/s, wall seconds of the slowest worker, dofs per worker)."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(procs)
    try:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [sample_n] * procs)
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    nd = res[0][1]
    return procs * nd / wall / 1e6, wall, nd


def host_cores(cap=64):
    """Cores this process may run on, capped (one oracle worker holds ~0.5 GB)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, cap))


def newton_step(n, deg, tile):
    """Newton-step wall time (BASELINE metric, second half): setup_solver() + solve() of the same MMS problem on an
    n x n mesh with the reference's default option set ('lu': served by GMRES + block-Jacobi/ILU(0) + coarse
    level run to 1e-12); per-step time = solve time / Newton iterations."""
    import torch
    import product_cases
    from echemfem_b200.mesh import UnitSquareMesh
    s = product_cases.AdvectionDiffusionMigrationSolver(n, 1.0, D1=0.5e-5, D2=1.0e-5, p=deg,
                                                        mesh=UnitSquareMesh(n, n, tile=tile if n % tile == 0 else None))
    s.init_solver_parameters(pc_type="lu")
    s.solver_parameters.pop("snes_monitor", None)
    s.setup_solver()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.solve()
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    e = s._engine
    its = max(e.last["snes_its"], 1)
    return {"mesh": "UnitSquareMesh(%d,%d) DG Q%d" % (n, n, deg), "dofs": e.ndof, "snes_its": e.last["snes_its"],
            "ksp_its_per_step": e.last["ksp_its_per_step"], "solve_s": t, "newton_step_s": t / its,
            "timers_s": {k: round(v, 4) for k, v in e.timers.items()},
            "linear_solver": ("GMRES(200) + geometric multigrid V-cycle (block-Jacobi/ILU(0) smoothers, Galerkin coarse "
                              "operators), ksp_rtol 1e-12" if getattr(e, "_mg", None) is not None else
                              "GMRES(200) + block-Jacobi/ILU(0) on %dx%d-cell tiles + aggregation coarse level, "
                              "ksp_rtol 1e-12" % (tile, tile))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample per step: K steps of `procs` concurrent 48 x 48 meshes end within a few minutes
    n = args.sample_n if args.sample_n != 96 else 48
    procs = args.cpu_procs or host_cores()
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    pool = mp.get_context("fork").Pool(procs)
    vals = []
    for _ in range(args.warmup):
        cpu_baseline_parallel(n, procs, pool)
    t0 = time.perf_counter()
    ndof = 0
    for _ in range(args.steps):
        v, t, nd = cpu_baseline_parallel(n, procs, pool)
        ndof = nd * procs
        vals.append(t)
    total = time.perf_counter() - t0
    pool.close()
    pool.join()
    value = ndof * args.steps / sum(vals) / 1e6
    sample = ("%d processes x UnitSquareMesh(%d,%d) DG Q1 = %d dofs per step (numpy oracle, complex-step Jacobian)"
              % (procs, n, n, ndof))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(vals) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2 MMS 2D DG Q1 electroneutral Nernst-Planck (bounded CPU sample)",
                       "sample_N": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": total}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--N", type=int, default=0, help="cells per direction (cfg2: 1408 / 944 / 704 for Q1 / Q2 / Q3)")
    ap.add_argument("--degree", type=int, default=1, help="polynomial degree (the headline number is Q1)")
    ap.add_argument("--sample-n", dest="sample_n", type=int, default=96, help="CPU baseline sample mesh")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--host-chunks", dest="host_chunks", type=int, default=16, help="cell ranges of the e2e "
                    "host-buffer pipeline (ech_set_host_chunks)")
    ap.add_argument("--cpu-procs", dest="cpu_procs", type=int, default=0, help="processes of the CPU baseline "
                                                                               "(default: every core of the host)")
    ap.add_argument("--tile", type=int, default=0, help="number the cells tile by tile (T x T cells) instead of "
                                                        "lexicographically")
    ap.add_argument("--extras", action="store_true", help="also time SpMV and one full Newton solve")
    ap.add_argument("--newton-n", dest="newton_n", type=int, default=1408, help="mesh of the --extras Newton solve "
                                                                               "(default: the cfg2 mesh itself)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries exactly one JSON line; everything the solver prints goes to stderr
    real_stdout = sys.stdout
    sys.stdout = sys.stderr

    import numpy as np
    import torch
    import ctypes as C
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import product_cases
    from echemfem_b200.newton import NewtonEngine
    from echemfem_b200 import capi
    from echemfem_b200.capi import ptr

    W = max(args.warmup, 3)
    N = args.N or {1: 1408, 2: 944, 3: 704}[args.degree]
    deg = args.degree
    # The path shards by cells.  Weak scaling: the global mesh is (world*N) x N cells on [0, world] x [0, 1],
    # rank r owns the r-th strip of N x N cells plus one ghost column per cut; per step the ghost-cell
    # state is exchanged (NCCL send/recv) before the residual and Jacobian kernels run.
    if world > 1:
        from echemfem_b200.partition import strip_partition
        lmesh, plan = strip_partition([np.linspace(0.0, float(world), world * N + 1), np.linspace(0.0, 1.0, N + 1)],
                                      rank, world)
        s = product_cases.AdvectionDiffusionMigrationSolver(0, 1.0, D1=0.5e-5, D2=1.0e-5, mesh=lmesh, p=deg)
    else:
        tmesh = None
        if args.tile:
            from echemfem_b200.mesh import UnitSquareMesh
            tmesh = UnitSquareMesh(N, N, tile=args.tile)
        s = product_cases.AdvectionDiffusionMigrationSolver(N, 1.0, D1=0.5e-5, D2=1.0e-5, p=deg, mesh=tmesh)
    eng = NewtonEngine(s)
    mesh = s.mesh
    x = s.V.node_coords()
    pert = 1e-3 * np.sin(7 * x[:, 0] + 3 * x[:, 1])
    u_host = np.concatenate([np.cos(x[:, 0]) + np.sin(x[:, 1]) + 3 + pert, np.sin(x[:, 0]) + np.cos(x[:, 1]) + 3 + pert])
    u = s.u.tensor
    u.copy_(torch.from_numpy(u_host).cuda())
    aux = eng._aux_tensor()
    consts = eng._consts()
    lib = eng.lib
    st = eng._stream()
    ndof = eng.nf * eng.ncell * eng.nloc            # owned dofs of this rank
    nstate = eng.ndof                               # owned + ghost

    def step():
        # one Newton-iteration assembly: F(u) and J(u) at the same state, one fused pass
        if eng.halo is not None:
            eng.halo.exchange(u)
        capi.check(lib.ech_residual_jacobian(eng.h, ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr), ptr(eng.vals),
                                             ptr(eng.F), st), eng.h)

    for _ in range(W):
        step()
    torch.cuda.synchronize()

    # per-kernel timings (events on the launching stream)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.ech_launch_count(eng.h)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    t_wall = time.perf_counter()
    for k in range(args.steps):
        evs[k][0].record()
        step()
        evs[k][1].record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall
    sampler.mark_end()
    clocks = sampler.stop()
    launches = lib.ech_launch_count(eng.h) - launches0
    t_step = evs[0][0].elapsed_time(evs[-1][1]) / args.steps
    # the two callbacks separately (what PETSc would call one after the other), for the breakdown
    sep = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tF = tJ = 0.0
    for _ in range(3):      # first launches of the two separate kernels (module load, attribute set-up) are not timed
        capi.check(lib.ech_residual(eng.h, ptr(u), ptr(aux), ptr(consts), ptr(eng.F), st), eng.h)
        capi.check(lib.ech_jacobian(eng.h, ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr), ptr(eng.vals), st), eng.h)
    torch.cuda.synchronize()
    for k in range(args.steps):
        sep[0].record()
        capi.check(lib.ech_residual(eng.h, ptr(u), ptr(aux), ptr(consts), ptr(eng.F), st), eng.h)
        sep[1].record()
        capi.check(lib.ech_jacobian(eng.h, ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr), ptr(eng.vals), st), eng.h)
        sep[2].record()
        torch.cuda.synchronize()
        tF += sep[0].elapsed_time(sep[1]) / args.steps
        tJ += sep[1].elapsed_time(sep[2]) / args.steps

    # end to end through the host-buffer C ABI: pinned host state in, residual out, every step
    uh = torch.from_numpy(u_host).pin_memory()
    capi.check(lib.ech_set_host_chunks(eng.h, args.host_chunks), eng.h)
    Fh = torch.empty(nstate, dtype=torch.float64).pin_memory()
    for _ in range(2):
        capi.check(lib.ech_assemble_host(eng.h, ptr(uh), ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr),
                                         ptr(eng.vals), ptr(eng.F), ptr(Fh), st), eng.h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        capi.check(lib.ech_assemble_host(eng.h, ptr(uh), ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr),
                                         ptr(eng.vals), ptr(eng.F), ptr(Fh), st), eng.h)
    e1.record()
    torch.cuda.synchronize()
    t_e2e = e0.elapsed_time(e1) / args.steps
    # the same call without the chunked pipeline (one copy in, one launch, one copy out), for the breakdown
    capi.check(lib.ech_set_host_chunks(eng.h, 0), eng.h)
    for _ in range(2):
        capi.check(lib.ech_assemble_host(eng.h, ptr(uh), ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr),
                                         ptr(eng.vals), ptr(eng.F), ptr(Fh), st), eng.h)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(min(args.steps, 10)):
        capi.check(lib.ech_assemble_host(eng.h, ptr(uh), ptr(u), ptr(aux), ptr(consts), ptr(eng.rowptr),
                                         ptr(eng.vals), ptr(eng.F), ptr(Fh), st), eng.h)
    e1.record()
    torch.cuda.synchronize()
    t_e2e_serial = e0.elapsed_time(e1) / min(args.steps, 10)
    capi.check(lib.ech_set_host_chunks(eng.h, args.host_chunks), eng.h)

    extras = {}
    if args.extras:
        xv = torch.from_numpy(np.cos(np.arange(nstate) * 0.37)).cuda()
        yv = torch.empty_like(xv)
        for _ in range(3):
            eng.spmv(xv, yv)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            eng.spmv(xv, yv)
        a1.record()
        torch.cuda.synchronize()
        extras["spmv_ms"] = a0.elapsed_time(a1) / args.steps
        if world == 1:
            del xv, yv
            torch.cuda.empty_cache()
            extras["newton"] = newton_step(args.newton_n, deg, args.tile or 8)

    times = torch.tensor([t_step, tF, tJ, t_e2e], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_step, tF, tJ, t_e2e = [float(v) for v in times.cpu()]

    if rank == 0:
        BJ, BF, BS = algorithmic_bytes(mesh, eng.nf, eng.nloc, eng.nnz)
        peak, peak_src = measured_peak()
        achieved = (BJ + 8 * ndof) / (t_step * 1e-3) / 1e9      # fused kernel: Jacobian bytes + the residual vector
        line = {
            "metric": METRIC, "value": world * ndof / (t_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": t_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2 MMS 2D DG Q%d electroneutral Nernst-Planck, UnitSquareMesh(%d,%d) quads%s" % (
                           deg, N, N, ", cells numbered in %dx%d tiles" % (args.tile, args.tile) if args.tile else ""),
                       "cells_per_gpu": getattr(mesh, "num_owned", mesh.num_cells), "dofs_per_gpu": ndof, "nnz_per_gpu": eng.nnz,
                       "l2_policy": "outputs (%.1f GB CSR values) exceed the 126 MB L2 every step" % (8 * eng.nnz / 1e9),
                       "parallelism": "cells partitioned in strips, one per GPU; ghost-cell state exchanged by NCCL "
                                      "send/recv every step (%d B per rank); no matrix entries communicated" % (
                                          eng.halo.bytes_per_exchange if eng.halo is not None else 0)},
            "residual_ms": tF, "jacobian_ms": tJ,
            "residual_MDOFs": ndof / (tF * 1e-3) / 1e6, "jacobian_MDOFs": ndof / (tJ * 1e-3) / 1e6,
            "roofline": {"bound": "hbm", "kernel": "echdg::v2::jacobian_coop2_kernel<true> (fused F + J, 1 launch per step)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "algorithmic_bytes": BJ + 8 * ndof, "traffic": ncu_traffic(deg, N),
                         "jacobian_only_achieved_GBs": BJ / (tJ * 1e-3) / 1e9,
                         "residual_achieved_GBs": BF / (tF * 1e-3) / 1e9},
            "e2e": {"value": world * ndof / (t_e2e * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": t_e2e,
                    "h2d_bytes_per_step": 8 * nstate, "d2h_bytes_per_step": 8 * nstate,
                    "host_chunks": args.host_chunks, "unchunked_ms_per_step": t_e2e_serial,
                    "path": "ech_assemble_host: pinned host state in, fused kernel per cell range, residual out; "
                            "copies of neighbouring ranges overlap the kernels on three streams"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
            "module_compile_s": eng.t_compile,
        }
        line.update(extras)
        if extras.get("spmv_ms"):
            line["spmv_GBs"] = BS / (extras["spmv_ms"] * 1e-3) / 1e9
        if not args.no_cpu:
            procs = args.cpu_procs or host_cores()
            v, t, nd = cpu_baseline_parallel(args.sample_n, procs)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
                                    "sample": "%d processes x UnitSquareMesh(%d,%d) DG Q1 (%d dofs each), one residual + "
                                              "one Jacobian (numpy oracle, complex-step Jacobian), %.1f s"
                                              % (procs, args.sample_n, args.sample_n, nd, t)}
        print(json.dumps(line), file=real_stdout, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
