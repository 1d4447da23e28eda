"""ctypes binding of libechem_b200.so (include/echem_b200.h).

The product path has no CPU fallback: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"`), importing this module's
`lib()` raises.
"""
import ctypes as C
import os

from . import build as _build

c_i64 = C.c_int64
c_i32 = C.c_int32
c_p = C.c_void_p


class MeshDesc(C.Structure):
    _fields_ = [("dim", c_i32), ("degree", c_i32), ("nfields", c_i32), ("naux", c_i32),
                ("ncell", c_i64), ("ncell_total", c_i64),
                ("xcell", c_p), ("nbr", c_p), ("code", c_p), ("slot", c_p), ("slot_self", c_p), ("ncol", c_p),
                ("cell_volume", c_p), ("cell_diam", c_p), ("facet_area", c_p),
                ("nnodes", c_i64), ("cell_nodes", c_p), ("inc_ptr", c_p), ("inc_cell", c_p), ("inc_loc", c_p),
                ("colslot", c_p), ("adj_count", c_p)]


class GmresOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("restart", c_i32), ("max_it", c_i32),
                ("precond", c_i32), ("monitor", c_i32), ("nfields", c_i32), ("nloc", c_i32),
                ("rows_per_field", c_i64), ("nblocks", c_i64), ("block_cell_ptr", c_p), ("lu", c_p),
                ("diag_pos", c_p), ("agg", c_p), ("nc", c_i32), ("reserved", c_i32), ("coarse_inv", c_p),
                ("coarse_work", c_p), ("lcol", c_p), ("passptr", c_p), ("slots", c_p), ("lanes_per_row", c_i32),
                ("max_block_rows", c_i32), ("refine", c_i32), ("reserved2", c_i32)]


EXPORTS = {
    "ech_create": (C.c_int, [C.c_char_p, C.POINTER(MeshDesc), C.POINTER(c_p)]),
    "ech_destroy": (None, [c_p]),
    "ech_last_error": (C.c_char_p, [c_p]),
    "ech_describe": (C.c_int, [c_p, C.POINTER(c_i32), C.c_int]),
    "ech_cell_geometry": (C.c_int, [c_i32, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "ech_num_rows": (C.c_int, [c_p, C.POINTER(c_i64)]),
    "ech_pattern_rowptr": (C.c_int, [c_p, c_p, C.POINTER(c_i64), c_p]),
    "ech_pattern_colidx": (C.c_int, [c_p, c_p, c_p, c_p]),
    "ech_residual": (C.c_int, [c_p, c_p, c_p, c_p, c_p, c_p]),
    "ech_jacobian": (C.c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ech_residual_jacobian": (C.c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ech_constrain_rows": (C.c_int, [c_i64, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ech_spmv": (C.c_int, [c_i64, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ech_axpby": (C.c_int, [c_i64, C.c_double, c_p, C.c_double, c_p, c_p]),
    "ech_dot": (C.c_int, [c_i64, c_p, c_p, c_p, C.POINTER(C.c_double), c_p]),
    "ech_multidot": (C.c_int, [c_i64, c_i32, c_p, c_p, c_p, c_p, c_p]),
    "ech_multiaxpy": (C.c_int, [c_i64, c_i32, c_p, c_p, C.c_double, c_p, c_p]),
    "ech_multiaxpy_dot": (C.c_int, [c_i64, c_i32, c_p, c_p, C.c_double, c_p, c_p, c_p, c_p]),
    "ech_bjacobi_ilu0_factor": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ech_bjacobi_ilu0_factor_async": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_p, c_p, c_p, c_i64, c_p, c_p,
                                                c_p, c_p]),
    "ech_bjacobi_ilu0_factor_status": (C.c_int, [c_i32]),
    "ech_bjacobi_ilu0_factor_planned": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_i32, c_p, c_p, c_p, c_i64,
                                                  c_p, c_p, c_p, c_p, c_p]),
    "ech_bjacobi_ilu0_apply": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ech_bjacobi_ilu0_plan_count": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_i32, c_p, c_p, c_i32, c_p, c_p,
                                              c_p, C.POINTER(c_i64), c_p]),
    "ech_bjacobi_ilu0_plan_fill": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_i32, c_p, c_p, c_i32, c_p, c_p,
                                             c_p, c_p]),
    "ech_bjacobi_ilu0_apply_planned": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_i32, c_p, c_p, c_p, c_p, c_p,
                                                 c_i32, c_p, c_p, c_p]),
    "ech_bjacobi_ilu0_pack_count": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_i32, c_p, c_p, c_i32, c_p, c_p,
                                              c_p, C.POINTER(c_i64), c_p]),
    "ech_bjacobi_ilu0_pack_fill": (C.c_int, [c_i64, c_i32, c_i64, c_i32, c_i64, c_p, c_i32, c_p, c_p, c_i32, c_p, c_p,
                                             c_p, c_p, c_p, c_p, c_p]),
    "ech_bjacobi_ilu0_pack_values": (C.c_int, [c_i64, c_i32, c_p, c_p, c_p, c_p, c_p]),
    "ech_bjacobi_ilu0_apply_packed": (C.c_int, [c_i32, c_i64, c_i32, c_i64, c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p,
                                                c_p]),
    "ech_halo_pack": (C.c_int, [c_i64, c_p, c_p, c_p, c_p]),
    "ech_halo_unpack": (C.c_int, [c_i64, c_p, c_p, c_p, c_p]),
    "ech_dg_prolong_add": (C.c_int, [c_i32, c_i32, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "ech_dg_restrict": (C.c_int, [c_i32, c_i32, c_i32, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "ech_aux_prolong_add": (C.c_int, [c_i32, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "ech_aux_galerkin": (C.c_int, [c_i32, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_p]),
    "ech_aux_galerkin_global": (C.c_int, [c_i32, c_i32, c_i64, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_i32, c_p, c_p, c_p,
                                          c_i64, c_p, c_p]),
    "ech_galerkin_dg": (C.c_int, [C.POINTER(c_i64), C.POINTER(c_p), c_p]),
    "ech_coarse_galerkin": (C.c_int, [c_i64, c_p, c_p, c_p, c_p, c_i32, c_p, c_p]),
    "ech_coarse_restrict": (C.c_int, [c_i64, c_p, c_i32, c_p, c_p, c_p]),
    "ech_coarse_prolong_add": (C.c_int, [c_i64, c_p, c_p, c_p, c_p]),
    "ech_dense_gemv": (C.c_int, [c_i32, c_i32, c_p, c_p, c_p, c_p]),
    "ech_gmres": (C.c_int, [c_i64, c_p, c_p, c_p, c_p, c_p, C.POINTER(GmresOpts), c_p, c_p,
                            C.POINTER(c_i32), C.POINTER(C.c_double), c_p]),
    "ech_assemble_host": (C.c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "ech_set_host_chunks": (C.c_int, [c_p, C.c_int]),
    "ech_launch_count": (c_i64, [c_p]),
}

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = _build.CORE_LIB
        if not os.path.exists(path):
            raise RuntimeError("libechem_b200.so is not built (%s); run __graft_entry__.build()" % path)
        L = C.CDLL(path)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class EchError(RuntimeError):
    pass


def check(rc, handle=None):
    if rc != 0:
        msg = lib().ech_last_error(handle)
        raise EchError("echem_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def ptr(t):
    """Raw device (or host) pointer of a torch tensor / None."""
    return None if t is None else C.c_void_p(t.data_ptr())
