"""ctypes binding of libmupopp_b200.so (the C-ABI declared in include/mupopp_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an
exception is raised.  PyTorch is used only to own device buffers and streams.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmupopp_b200.so")

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class MpError(RuntimeError):
    pass


class MpMesh(C.Structure):
    _fields_ = [("dim", C.c_int32), ("ncell", C.c_int64), ("nvert", C.c_int64),
                ("d_cells", C.c_void_p), ("d_coords", C.c_void_p), ("d_cell_dofs", C.c_void_p),
                ("d_row_pos", C.c_void_p)]


class MpTransportParams(C.Structure):
    _fields_ = [("Pe", C.c_double), ("Da", C.c_double), ("phi", C.c_double), ("beta", C.c_double), ("dt", C.c_double),
                ("sigma", C.c_double), ("rho0", C.c_double), ("gamma", C.c_double),
                ("frac", C.c_double * 3), ("zhat", C.c_double * 3), ("Kconst", C.c_double),
                ("reaction", C.c_int32), ("ncomp", C.c_int32), ("supg", C.c_int32), ("pad", C.c_int32)]


class MpQuad(C.Structure):
    _fields_ = [("nq", C.c_int32), ("nbv", C.c_int32), ("d_w", C.c_void_p), ("d_lam", C.c_void_p), ("d_vphi", C.c_void_p),
                ("d_vdphi", C.c_void_p)]


class MpCompactionParams(C.Structure):
    _fields_ = [("da", C.c_double), ("R", C.c_double), ("B", C.c_double), ("theta", C.c_double), ("dL", C.c_double),
                ("zhat", C.c_double * 3)]


class MpFunctional(C.Structure):
    _fields_ = [("nfield", C.c_int32), ("nb", C.c_int32 * 4), ("d_field", C.c_void_p * 4), ("d_dofs", C.c_void_p * 4),
                ("d_tab", C.c_void_p * 4), ("vec_kind", C.c_int32), ("nbv", C.c_int32), ("d_vec", C.c_void_p * 3),
                ("d_vdofs", C.c_void_p), ("d_vtab", C.c_void_p), ("nq", C.c_int32), ("ntab", C.c_int32), ("d_w", C.c_void_p)]


class MpCsr(C.Structure):
    _fields_ = [("nrow", C.c_int64), ("ncol", C.c_int64), ("nnz", C.c_int64), ("max_row_nnz", C.c_int32), ("pad", C.c_int32),
                ("d_rowptr", C.c_void_p), ("d_colidx", C.c_void_p), ("d_vals", C.c_void_p),
                ("d_sell_sliceptr", C.c_void_p), ("d_sell_col", C.c_void_p), ("d_sell_val", C.c_void_p),
                ("d_sell_meta", C.c_void_p), ("d_sell_xcol", C.c_void_p)]


class MpBlockTerm(C.Structure):
    _fields_ = [("rb", C.c_int32), ("cb", C.c_int32), ("A", C.POINTER(MpCsr)), ("scale", C.c_double)]


class MpBlockOp(C.Structure):
    _fields_ = [("nblk", C.c_int32), ("nterm", C.c_int32), ("off", C.c_int64 * 9), ("terms", C.POINTER(MpBlockTerm)), ("d_mask", C.c_void_p)]


class MpBlockPrec(C.Structure):
    _fields_ = [("P", C.POINTER(MpCsr) * 8), ("degree", C.c_int32 * 8), ("lmin", C.c_double * 8), ("lmax", C.c_double * 8),
                ("d_dinv", C.c_void_p)]


class MpSolveInfo(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("maxit", C.c_int32), ("check_every", C.c_int32),
                ("niter", C.c_int32), ("converged", C.c_int32), ("relres", C.c_double)]


# name -> (restype, argtypes); every name here must be declared in include/mupopp_b200.h
_VP = C.c_void_p
SIGNATURES = {
    "mp_version": (C.c_int, []),
    "mp_last_error": (C.c_char_p, []),
    "mp_ctx_create": (C.c_int, [C.c_int, _VP, C.POINTER(_VP)]),
    "mp_ctx_destroy": (C.c_int, [_VP]),
    "mp_ctx_sync": (C.c_int, [_VP]),
    "mp_ctx_launch_count": (C.c_int64, [_VP]),
    "mp_ctx_timing": (C.c_int, [_VP, C.c_int, C.c_int]),
    "mp_ctx_timing_read": (C.c_int, [_VP, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "mp_plan_create": (C.c_int, [_VP, _VP, C.c_int32, _VP, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_VP)]),
    "mp_plan_nnz": (C.c_int64, [_VP]),
    "mp_plan_max_row_nnz": (C.c_int32, [_VP]),
    "mp_ctx_set_option": (C.c_int, [_VP, C.c_char_p, C.c_int]),
    "mp_plan_export_csr": (C.c_int, [_VP, _VP, _VP, _VP]),
    "mp_plan_destroy": (C.c_int, [_VP]),
    "mp_plan_sell_info": (C.c_int, [_VP, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(_VP), C.POINTER(_VP)]),
    "mp_plan_sellu_info": (C.c_int, [_VP, C.POINTER(C.c_int64), C.POINTER(_VP), C.POINTER(_VP)]),
    "mp_csr_to_sell": (C.c_int, [_VP, _VP, _VP, _VP]),
    "mp_scatter_matrix": (C.c_int, [_VP, _VP, _VP, _VP]),
    "mp_scatter_matrix_ordered": (C.c_int, [_VP, _VP, _VP, _VP]),
    "mp_scatter_vector_ordered": (C.c_int, [_VP, _VP, _VP, C.c_double, _VP]),
    "mp_plan_row_positions": (C.c_int, [_VP, C.POINTER(_VP)]),
    "mp_scatter_vector": (C.c_int, [_VP, _VP, _VP, C.c_double, _VP]),
    "mp_apply_dirichlet_sym": (C.c_int, [_VP, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mp_elem_p1_mass": (C.c_int, [_VP, C.POINTER(MpMesh), C.c_double, _VP]),
    "mp_elem_p1_stiffness": (C.c_int, [_VP, C.POINTER(MpMesh), _VP, C.c_double, C.c_double, _VP]),
    "mp_elem_p1_reaction_rhs": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpTransportParams), _VP, _VP, _VP, _VP, _VP]),
    "mp_elem_p1_transport": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpTransportParams), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mp_elem_p1_adr_single": (C.c_int, [_VP, C.POINTER(MpMesh), C.c_double, C.c_double, C.c_double, _VP, _VP, _VP, _VP]),
    "mp_elem_p1_pressure_rhs": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpTransportParams), _VP, _VP, _VP]),
    "mp_recover_velocity": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpTransportParams), _VP, _VP, _VP, _VP]),
    "mp_elem_p1_cell_load": (C.c_int, [_VP, C.POINTER(MpMesh), _VP, C.c_int, _VP]),
    "mp_elem_reftensor": (C.c_int, [_VP, C.POINTER(MpMesh), C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, C.c_int, C.c_double, _VP]),
    "mp_elem_reftensor_rows": (C.c_int, [_VP, C.POINTER(MpMesh), C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, C.c_int, C.c_double, _VP, _VP]),
    "mp_elem_p1_transport_quad": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpTransportParams), C.POINTER(MpQuad), _VP, _VP, _VP, _VP,
                                            _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mp_cell_fn": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpQuad), C.c_int, C.c_double, _VP, _VP]),
    "mp_elem_p1_mass_fn": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpQuad), C.c_int, C.c_double, C.c_double, _VP, _VP]),
    "mp_elem_compaction_rhs": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpQuad), C.POINTER(MpCompactionParams), _VP, _VP, _VP,
                                         _VP, _VP, _VP, _VP]),
    "mp_elem_p1_porosity_quad": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpQuad), C.c_double, C.c_double, _VP, _VP, _VP, _VP,
                                           _VP, _VP, _VP, _VP]),
    "mp_elem_p2_porosity_quad": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpQuad), C.c_double, C.c_double, _VP, _VP, _VP, _VP,
                                           _VP, _VP, _VP, _VP]),
    "mp_integrate": (C.c_int, [_VP, C.c_int32, C.POINTER(MpFunctional), C.c_int64, _VP, _VP, _VP, _VP, C.POINTER(C.c_double),
                               C.POINTER(C.c_double)]),
    "mp_max_cell_speed": (C.c_int, [_VP, C.c_int32, C.c_int64, C.c_int32, _VP, _VP, _VP, _VP, C.c_int32, _VP, C.c_int32, _VP,
                                    C.POINTER(C.c_double)]),
    "mp_spmv": (C.c_int, [_VP, C.POINTER(MpCsr), _VP, _VP]),
    "mp_dot": (C.c_int, [_VP, C.c_int64, _VP, _VP, C.POINTER(C.c_double)]),
    "mp_axpby": (C.c_int, [_VP, C.c_int64, C.c_double, _VP, C.c_double, _VP]),
    "mp_vmul": (C.c_int, [_VP, C.c_int64, C.c_double, _VP, _VP, _VP]),
    "mp_jacobi_setup": (C.c_int, [_VP, C.POINTER(MpCsr), _VP]),
    "mp_solve_pcg": (C.c_int, [_VP, C.POINTER(MpCsr), _VP, _VP, _VP, _VP, C.POINTER(MpSolveInfo)]),
    "mp_spectral_bound": (C.c_int, [_VP, C.POINTER(MpCsr), _VP, C.POINTER(C.c_double)]),
    "mp_solve_pcg_chebyshev": (C.c_int, [_VP, C.POINTER(MpCsr), _VP, _VP, _VP, _VP, C.c_int, C.c_double, C.c_double, C.POINTER(MpSolveInfo)]),
    "mp_plan_incidence": (C.c_int, [_VP, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    "mp_assemble_p1_transport_rows": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpTransportParams), _VP] + [_VP] * 7 + [_VP, _VP]
                                      + [_VP, _VP, _VP, _VP, C.c_int64, _VP, _VP, _VP]),
    "mp_line_factor": (C.c_int, [_VP, C.c_int64, C.c_int64, C.c_int, _VP, _VP, _VP]),
    "mp_elem_p1_reaction_load": (C.c_int, [_VP, C.POINTER(MpMesh), C.POINTER(MpTransportParams), _VP, _VP, C.c_double, _VP]),
    "mp_line_setup": (C.c_int, [_VP, C.POINTER(MpCsr), C.c_int64, C.c_int, _VP, _VP, _VP]),
    "mp_solve_bicgstab_line": (C.c_int, [_VP, C.POINTER(MpCsr), C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, C.POINTER(MpSolveInfo)]),
    "mp_line_apply": (C.c_int, [_VP, C.c_int64, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mp_solve_pcg_line": (C.c_int, [_VP, C.POINTER(MpCsr), C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, C.POINTER(MpSolveInfo)]),
    "mp_solve_bicgstab": (C.c_int, [_VP, C.POINTER(MpCsr), _VP, _VP, _VP, _VP, C.POINTER(MpSolveInfo)]),
    "mp_solve_gmres": (C.c_int, [_VP, C.POINTER(MpCsr), _VP, _VP, _VP, _VP, C.c_int, C.POINTER(MpSolveInfo)]),
    "mp_block_apply": (C.c_int, [_VP, C.POINTER(MpBlockOp), _VP, _VP]),
    "mp_solve_minres": (C.c_int, [_VP, C.POINTER(MpBlockOp), C.POINTER(MpBlockPrec), _VP, _VP, C.POINTER(MpSolveInfo)]),
    "mp_schur_darcy_solve": (C.c_int, [_VP, C.c_int32, C.POINTER(MpCsr), C.POINTER(C.POINTER(MpCsr)), C.POINTER(C.POINTER(MpCsr)),
                                       C.POINTER(MpCsr), C.c_double, _VP, _VP, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double,
                                       C.c_double, _VP, _VP, C.POINTER(MpSolveInfo)]),
    "mp_nccl_unique_id": (C.c_int, [_VP]),
    "mp_comm_init": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "mp_comm_destroy": (C.c_int, [_VP]),
    "mp_allreduce_sum": (C.c_int, [_VP, _VP, C.c_int]),
    "mp_halo_create": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, _VP, _VP, C.POINTER(_VP)]),
    "mp_halo_exchange": (C.c_int, [_VP, _VP, _VP]),
    "mp_halo_destroy": (C.c_int, [_VP]),
    "mp_comm_peer_export": (C.c_int, [_VP, _VP]),
    "mp_comm_peer_attach": (C.c_int, [_VP, _VP]),
    "mp_comm_peer_enabled": (C.c_int, [_VP]),
    "mp_halo_peer_export": (C.c_int, [_VP, _VP, C.c_int64, C.c_int64, _VP]),
    "mp_halo_peer_attach": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises MpError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpError("%s not found: build it with `python -m mupopp_b200.build` (or __graft_entry__.build()); "
                      "mupopp_b200 has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mp_last_error()
        raise MpError("%s failed (code %d): %s" % (what or "mupopp_b200 call", rc, msg.decode() if msg else "?"))


def ptr(t):
    """Raw device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())
