"""CPU port of the DG assembly hot path (C++/OpenMP, `assemble_dg.cpp`) -- TEST / BASELINE INFRASTRUCTURE ONLY.

Used by `bench.py`'s `cpu_baseline` and `--impl reference` legs ("kind": "port": the reference itself --
Firedrake/PETSc -- is not installable offline) and by `tests/` as a second oracle at sizes the numpy oracle cannot
reach.  The product package never imports this.  Built with g++ (no CUDA), on first use, into `oracle/_build/`.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(os.path.dirname(HERE), "_build")
FLAGS = ["-O3", "-march=x86-64-v3", "-mtune=native", "-fopenmp", "-std=c++17", "-shared", "-fPIC"]


def _src_hash():
    h = hashlib.sha256()
    for name in ("assemble_dg.cpp", "host_shim.h"):
        with open(os.path.join(HERE, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()[:8]


def build(header_text, key):
    """Compile the port against one generated pointwise-integrand header; returns the library path (cached)."""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libcpu_port_%s_%s.so" % (key, _src_hash()))
    if os.path.exists(out):
        return out
    hdr = os.path.join(BUILD, "gen_%s.h" % key)
    with open(hdr, "w") as f:
        f.write(header_text)
    tmp = out + ".tmp%d" % os.getpid()
    cmd = ["g++"] + FLAGS + ["-include", os.path.join(HERE, "host_shim.h"), "-include", hdr,
                             os.path.join(HERE, "assemble_dg.cpp"), "-o", tmp]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cpu_port build failed:\n" + r.stderr[-4000:])
    os.replace(tmp, out)
    return out


class _Mesh(C.Structure):
    _fields_ = [("ncell", C.c_int64), ("ncell_total", C.c_int64), ("xcell", C.c_void_p), ("nbr", C.c_void_p),
                ("code", C.c_void_p), ("slot", C.c_void_p), ("slot_self", C.c_void_p), ("ncol", C.c_void_p),
                ("cell_volume", C.c_void_p), ("cell_diam", C.c_void_p), ("facet_area", C.c_void_p)]


def cell_geometry(xcell):
    """CellVolume, FacetArea, CellDiameter (echemfem/solver.py:1491, 1657) of multilinear cells
    given as (nc, 2^d, d) vertex coordinates: exact integrals of |J| (3-point Gauss), max vertex distance."""
    from .. import fem
    nc, nv, d = xcell.shape

    E = xcell[:, 1:] - xcell[:, :1]          # edge form: exact differences, no cancellation far from the origin

    def jac(xi):
        _, dN = fem.tabulate(np.array([0.0, 1.0]), xi)
        return np.einsum("qvb,eva->eqab", dN[:, 1:], E)
    xi, w = fem.cell_quadrature(3, d)
    vol = np.einsum("q,eq->e", w, np.abs(np.linalg.det(jac(xi))))
    area = np.zeros((nc, 2 * d))
    t, wf = fem.facet_quadrature(3, d)
    for lf in range(2 * d):
        J = jac(fem.facet_to_cell(lf, t, d))
        k = lf // 2
        g = np.linalg.inv(J)[..., k, :]
        area[:, lf] = np.einsum("q,eq->e", wf, np.abs(np.linalg.det(J)) * np.sqrt((g ** 2).sum(-1)))
    diff = xcell[:, :, None, :] - xcell[:, None, :, :]
    diam = np.sqrt((diff ** 2).sum(-1)).max(axis=(1, 2))
    return vol, area, diam


class CpuPort:
    """The port bound to one product solver (mesh + weak form).  Host arrays only; no CUDA anywhere."""

    def __init__(self, solver, nthreads=0):
        from echemfem_b200.compiler import CompiledForm
        mesh, W = solver.mesh, solver.W
        V = W.scalar
        if V.family != "DG":
            raise NotImplementedError("cpu_port: discontinuous spaces only")
        self.nf, self.naux, self.n = W.num_sub_spaces(), len(solver._aux), V.n
        self.form = CompiledForm(solver.Form, self.nf, self.naux, mesh.dim, V.degree, V.family)
        self.lib = C.CDLL(build(self.form.header, self.form.key))
        self.lib.cpu_port_assemble.restype = C.c_int
        self.lib.cpu_port_assemble.argtypes = [C.POINTER(_Mesh)] + [C.c_void_p] * 8 + [C.c_int64, C.c_int64, C.c_int]
        self.nthreads = int(nthreads)
        self.ncell_total = mesh.num_cells
        self.ncell = getattr(mesh, "num_owned", mesh.num_cells)
        nbr, code, slot, slot_self, ncol = mesh.connectivity(self.form.marker_class)
        c = np.ascontiguousarray
        self.xcell = c(mesh.cell_coords(), dtype=np.float64)
        self.nbr, self.code, self.slot = c(nbr, dtype=np.int32), c(code, dtype=np.uint8), c(slot, dtype=np.uint8)
        self.slot_self, self.ncol = c(slot_self, dtype=np.uint8), c(ncol, dtype=np.uint8)
        d = mesh.dim
        vol, area, diam = cell_geometry(self.xcell.reshape(self.ncell_total, 2 ** d, d))
        self.vol, self.area, self.diam = c(vol), c(area[:self.ncell]), c(diam)
        self.N = self.ncell_total * self.n
        self.ndof = self.nf * self.N
        self.consts = c(self.form.constant_values(), dtype=np.float64)
        self.aux = c(np.concatenate([fn.numpy() for fn in solver._aux])) if self.naux else None
        self.rowptr = self._rowptr()
        self.nnz = int(self.rowptr[-1])
        p = lambda a: None if a is None else a.ctypes.data
        self._mesh = _Mesh(self.ncell, self.ncell_total, p(self.xcell), p(self.nbr), p(self.code), p(self.slot),
                           p(self.slot_self), p(self.ncol), p(self.vol), p(self.diam), p(self.area))

    def _rowptr(self):
        """Row pointer of the field-blocked AIJ pattern: row (i, c, a) holds, for every coupled
        field j, `ncol[c]` cell blocks if the coupling crosses facets, else the own-cell block."""
        own, nb = self.form.own_mask, self.form.nbr_mask
        ncol = np.zeros(self.ncell_total, dtype=np.int64)
        ncol[:self.ncell] = self.ncol
        lens = []
        for i in range(self.nf):
            per_cell = sum((ncol if nb[i, j] else (ncol > 0).astype(np.int64)) * self.n
                           for j in range(self.nf) if own[i, j])
            lens.append(np.repeat(per_cell, self.n))
        return np.concatenate([[0], np.cumsum(np.concatenate(lens))]).astype(np.int64)

    def assemble(self, u, jacobian=True, residual=True, cells=None, scales=False, out=None):
        """(F, vals) at state u (field-blocked, owned + ghost layout); either may be None if not requested.
        scales: also return (F_scale, vals_scale), the sums of absolute element-level contributions to every entry
        (the tau_block of tests/parity.py).  out: (F, vals) arrays to write into (rows outside `cells` untouched)."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.size == self.ndof
        if out is not None:
            F, vals = out
        else:
            F = np.zeros(self.ndof) if residual else None
            vals = np.zeros(self.nnz) if jacobian else None
        Fs = np.zeros(self.ndof) if (scales and F is not None) else None
        vs = np.zeros(self.nnz) if (scales and vals is not None) else None
        c0, c1 = (0, self.ncell) if cells is None else cells
        p = lambda a: None if a is None else a.ctypes.data
        rc = self.lib.cpu_port_assemble(C.byref(self._mesh), p(u), p(self.aux), p(self.consts), p(self.rowptr),
                                        p(vals), p(F), p(vs), p(Fs), c0, c1, self.nthreads)
        assert rc == 0
        return (F, vals, Fs, vs) if scales else (F, vals)
