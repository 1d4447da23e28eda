"""Plan of the level-scheduled block-Jacobi / ILU(0) solves (`ech_bjacobi_ilu0_plan_*`, include/echem_b200.h).

PETSc's `bjacobi` + `sub_pc_type ilu` (echemfem/solver.py:1186-1188) applies L and U row by row; the plan sorts the
rows of every block by dependency level once per sparsity pattern, so that the solve walks wavefronts instead of rows.
Device buffers are torch tensors owned here; the library only fills them.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import capi
from .capi import ptr

MAX_BLOCK_ROWS = 16384


class IluPlan:
    def __init__(self, lib, dev, stream, ndof, nf, rows_per_field, unit, block_edges, block_ptr, rowptr, colidx, nnz,
                 pack=None):
        if pack is None:
            pack = os.environ.get("ECH_ILU_PACKED", "1") != "0"
        self.lib, self.ndof, self.nf, self.N, self.unit = lib, int(ndof), int(nf), int(rows_per_field), int(unit)
        self.nblocks = len(block_edges) - 1
        self.block_ptr = block_ptr
        self.max_rows = int(np.diff(np.asarray(block_edges)).max()) * self.unit * self.nf if self.nblocks else 0
        self.ok = 0 < self.max_rows <= MAX_BLOCK_ROWS and nnz > 0
        if not self.ok:
            return
        # lanes sharing a row hold 10 entries each in registers; a half-row has at most (row length - 1) entries
        row = nnz / max(self.ndof, 1)
        self.lpr = 4 if row <= 41 else (8 if row <= 81 else 16)
        G = 32 // self.lpr
        self.lcol = torch.empty(nnz, dtype=torch.int16, device=dev)
        levels = torch.empty(2 * self.ndof, dtype=torch.int16, device=dev)
        self.passptr = torch.zeros(2 * (self.nblocks + 1), dtype=torch.int64, device=dev)
        npass = C.c_int64()
        capi.check(lib.ech_bjacobi_ilu0_plan_count(self.ndof, self.nf, self.N, self.unit, self.nblocks, ptr(block_ptr),
                                                   self.max_rows, ptr(rowptr), ptr(colidx), self.lpr, ptr(self.lcol),
                                                   ptr(levels), ptr(self.passptr), C.byref(npass), stream))
        self.npass = int(npass.value)
        self.slots = torch.empty(max(self.npass * G, 1), dtype=torch.int64, device=dev)
        capi.check(lib.ech_bjacobi_ilu0_plan_fill(self.ndof, self.nf, self.N, self.unit, self.nblocks, ptr(block_ptr),
                                                  self.max_rows, ptr(rowptr), ptr(colidx), self.lpr, ptr(levels),
                                                  ptr(self.passptr), ptr(self.slots), stream))
        self.rowptr = rowptr
        self.packed = None
        if pack:
            self._pack_plan(levels, rowptr, colidx, block_ptr, stream, dev, nnz)
        del levels

    # -- pass-packed factors fetched with bulk copies (ech_bjacobi_ilu0_pack_*) ---------------------------------
    def _pack_plan(self, levels, rowptr, colidx, block_ptr, stream, dev, nnz):
        lib = self.lib
        sizes = torch.zeros(6 * (self.nblocks + 1), dtype=torch.int64, device=dev)
        tot = (C.c_int64 * 4)()
        capi.check(lib.ech_bjacobi_ilu0_pack_count(self.ndof, self.nf, self.N, self.unit, self.nblocks, ptr(block_ptr),
                                                   self.max_rows, ptr(rowptr), ptr(colidx), self.lpr, ptr(levels),
                                                   ptr(self.lcol), ptr(sizes), tot, stream))
        nbytes, nent, npass, maxchunk = int(tot[0]), int(tot[1]), int(tot[2]), int(tot[3])
        ring_stride = (maxchunk + 15) // 16 * 16
        per_warp = (self.max_rows * 8 + 15) // 16 * 16 + 64 + 4 * ring_stride
        if nbytes == 0 or per_warp * 4 > 227 * 1024 or nent >= 2 ** 32:
            return                                      # blocks too large for the ring: the planned solve stays
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.src = torch.empty(max(nent, 1), dtype=torch.int32, device=dev)
        self.bdesc = torch.zeros(2 * self.nblocks * 8, dtype=torch.int64, device=dev)
        self.maxchunk = maxchunk
        capi.check(lib.ech_bjacobi_ilu0_pack_fill(self.ndof, self.nf, self.N, self.unit, self.nblocks, ptr(block_ptr),
                                                  self.max_rows, ptr(rowptr), ptr(colidx), self.lpr, ptr(levels),
                                                  ptr(self.lcol), ptr(sizes), ptr(self.packed), ptr(self.src),
                                                  ptr(self.bdesc), stream))
        self._packed_for = None

    def factor_async(self, rowptr, colidx, vals, lu, diag, flag_ptr, stream):
        """lu = ILU(0)(vals) block by block on the plan's local columns (`ech_bjacobi_ilu0_factor_planned`), then the
        packed copy of the factors; no host synchronisation: `flag_ptr` is a device int32 the caller reads later and
        hands to `ech_bjacobi_ilu0_factor_status`."""
        capi.check(self.lib.ech_bjacobi_ilu0_factor_planned(self.ndof, self.nf, self.N, self.unit, self.nblocks,
                                                            ptr(self.block_ptr), self.max_rows, ptr(rowptr), ptr(colidx),
                                                            ptr(vals), int(vals.numel()), ptr(self.lcol), ptr(lu),
                                                            ptr(diag), flag_ptr, stream))
        self.refresh(lu, stream)

    def factor(self, rowptr, colidx, vals, lu, diag, stream):
        """Synchronous form of `factor_async` (reads the status word back)."""
        if getattr(self, "_flag", None) is None:
            self._flag = torch.zeros(1, dtype=torch.int32, device=lu.device)
        self.factor_async(rowptr, colidx, vals, lu, diag, ptr(self._flag), stream)
        capi.check(self.lib.ech_bjacobi_ilu0_factor_status(int(self._flag.item())))

    def refresh(self, lu, stream):
        """After every factorisation: the values of `lu` into the packed chunks."""
        if self.packed is not None:
            capi.check(self.lib.ech_bjacobi_ilu0_pack_values(self.nblocks, self.lpr, ptr(self.bdesc), ptr(self.packed),
                                                             ptr(self.src), ptr(lu), stream))
            self._packed_for = lu.data_ptr()

    def apply(self, lu, r, z, stream):
        if self.packed is not None and self._packed_for == lu.data_ptr():
            capi.check(self.lib.ech_bjacobi_ilu0_apply_packed(self.nf, self.N, self.unit, self.nblocks,
                                                              ptr(self.block_ptr), self.max_rows, self.maxchunk,
                                                              self.lpr, ptr(self.bdesc), ptr(self.packed), ptr(r),
                                                              ptr(z), stream))
            return
        capi.check(self.lib.ech_bjacobi_ilu0_apply_planned(self.ndof, self.nf, self.N, self.unit, self.nblocks,
                                                           ptr(self.block_ptr), self.max_rows, ptr(self.rowptr),
                                                           ptr(lu), ptr(self.lcol), ptr(self.passptr), ptr(self.slots),
                                                           self.lpr, ptr(r), ptr(z), stream))

    def set_opts(self, opts):
        """Fill the plan fields of an `ech_gmres_opts`."""
        opts.lcol, opts.passptr, opts.slots = ptr(self.lcol), ptr(self.passptr), ptr(self.slots)
        opts.lanes_per_row, opts.max_block_rows = self.lpr, self.max_rows
