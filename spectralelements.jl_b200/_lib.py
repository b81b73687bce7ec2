"""ctypes binding of libsemb.so (include/semb.h).  No fallback: a missing library is an ImportError."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SEMB_LIB: an experimental variant built with `build.py --tag` (A/B tools only; the product library is lib/libsemb.so)
LIB_PATH = os.environ.get("SEMB_LIB") or os.path.join(HERE, "lib", "libsemb.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ll_p = C.POINTER(C.c_longlong)
vp = C.c_void_p


class PcgOpts(C.Structure):
    """struct semb_pcg_opts (include/semb.h)"""
    _fields_ = [
        ("nu", C.c_double), ("nu_arr", vp),
        ("k", C.c_double), ("k_arr", vp),
        ("bc", C.c_char_p), ("M_arr", vp),
        ("precond", C.c_int), ("prec_b0", C.c_double),
        ("tol", C.c_double), ("maxiter", C.c_longlong),
        ("check_every", C.c_int),
    ]


class SembError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsemb error %d: %s" % (code, msg))
        self.code = code


# name -> (argtypes, restype is always c_int unless given)
_SIGS = {
    "semb_version": ([], C.c_int),
    "semb_last_error": ([], C.c_char_p),
    "semb_init": ([C.c_int, C.POINTER(vp)], C.c_int),
    "semb_finalize": ([vp], C.c_int),
    "semb_sync": ([vp], C.c_int),
    "semb_stream": ([vp, C.POINTER(vp)], C.c_int),
    "semb_timer_start": ([vp], C.c_int),
    "semb_timer_stop": ([vp, c_double_p], C.c_int),
    "semb_launch_count": ([vp, c_ll_p], C.c_int),
    "semb_flush_l2": ([vp], C.c_int),
    "semb_profile_enable": ([vp, C.c_int], C.c_int),
    "semb_profile_read": ([vp, c_double_p, c_int_p], C.c_int),
    "semb_alloc_pinned": ([C.c_size_t, C.POINTER(vp)], C.c_int),
    "semb_free_pinned": ([vp], C.c_int),
    "semb_partition": ([C.c_int, C.c_int, C.c_int, c_int_p, c_int_p], C.c_int),
    "semb_plan_chunks": ([C.c_int, C.c_int, C.c_int, c_int_p], C.c_int),
    "semb_halo_plan": ([C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p], C.c_int),
    "semb_comm_unique_id": ([C.c_char_p], C.c_int),
    "semb_comm_init": ([vp, C.c_int, C.c_int, C.c_char_p], C.c_int),
    "semb_comm_info": ([vp, c_int_p, c_int_p], C.c_int),
    "semb_comm_barrier": ([vp], C.c_int),
    "semb_comm_allreduce_max": ([vp, c_double_p, C.c_int], C.c_int),
    "semb_gausslobatto": ([C.c_int, c_double_p, c_double_p], C.c_int),
    "semb_deriv_mat": ([C.c_int, c_double_p, c_double_p], C.c_int),
    "semb_interp_mat": ([C.c_int, c_double_p, C.c_int, c_double_p, c_double_p], C.c_int),
    "semb_semmesh": ([C.c_int, C.c_int, c_double_p, c_double_p], C.c_int),
    "semb_bdf_ext_k": ([C.c_int, c_double_p, C.c_int, c_double_p, c_double_p], C.c_int),
    "semb_mesh_create_xy": ([vp] + [C.c_int] * 6 + [c_double_p] * 6 + [C.POINTER(vp)], C.c_int),
    "semb_mesh_create_deform": ([vp] + [C.c_int] * 7 + [c_double_p, C.c_int, C.POINTER(vp)], C.c_int),
    "semb_mesh_create_arrays": ([vp] + [C.c_int] * 6 + [c_double_p] * 6 + [C.POINTER(vp)], C.c_int),
    "semb_mesh_destroy": ([vp], C.c_int),
    "semb_mesh_dims": ([vp] + [c_int_p] * 8, C.c_int),
    "semb_mesh_get": ([vp, C.c_int, c_double_p], C.c_int),
    "semb_mesh_set": ([vp, C.c_int, c_double_p], C.c_int),
    "semb_mesh_get_D": ([vp, c_double_p, c_double_p], C.c_int),
    "semb_generate_mask": ([vp, C.c_char_p, c_double_p], C.c_int),
    "semb_field_create": ([vp, C.POINTER(vp)], C.c_int),
    "semb_field_destroy": ([vp], C.c_int),
    "semb_field_upload": ([vp, c_double_p], C.c_int),
    "semb_field_download": ([vp, c_double_p], C.c_int),
    "semb_field_fill": ([vp, C.c_double], C.c_int),
    "semb_field_copy": ([vp, vp], C.c_int),
    "semb_field_fill_random": ([vp, C.c_uint64], C.c_int),
    "semb_field_axpby": ([C.c_double, vp, C.c_double, vp], C.c_int),
    "semb_field_devptr": ([vp, C.POINTER(vp), c_ll_p], C.c_int),
    "semb_lapl": ([vp, vp, vp], C.c_int),
    "semb_hlmz": ([vp, vp, vp, C.c_double, vp, C.c_double, vp], C.c_int),
    "semb_mass": ([vp, vp, vp], C.c_int),
    "semb_gather_scatter": ([vp, vp, vp], C.c_int),
    "semb_mask": ([vp, vp, vp, vp], C.c_int),
    "semb_mask_bc": ([vp, vp, C.c_char_p, vp], C.c_int),
    "semb_oplhs": ([vp, vp, vp, C.c_double, vp, C.c_double, C.c_char_p, vp, vp], C.c_int),
    "semb_jac": ([vp] * 9, C.c_int),
    "semb_dot_mult": ([vp, vp, vp, c_double_p], C.c_int),
    "semb_norm_inf": ([vp, vp, c_double_p], C.c_int),
    "semb_pcg": ([vp, C.POINTER(PcgOpts), vp, vp, c_ll_p, c_double_p], C.c_int),
    "semb_pcg_begin": ([vp, C.POINTER(PcgOpts), vp, vp], C.c_int),
    "semb_pcg_iterate": ([vp, C.c_int], C.c_int),
    "semb_pcg_status": ([vp, c_ll_p, c_double_p, c_int_p], C.c_int),
    "semb_diffusion_create": ([vp, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(vp)], C.c_int),
    "semb_diffusion_destroy": ([vp], C.c_int),
    "semb_diffusion_field": ([vp, C.c_int, C.POINTER(vp)], C.c_int),
    "semb_diffusion_begin_step": ([vp, c_double_p, c_ll_p], C.c_int),
    "semb_diffusion_finish_step": ([vp, C.c_double, c_ll_p, c_double_p], C.c_int),
    "semb_diffusion_state": ([vp, c_double_p, c_double_p, c_double_p, c_ll_p], C.c_int),
    "semb_diffusion_set_precond": ([vp, C.c_int], C.c_int),
    "semb_grad": ([vp, vp, vp, vp], C.c_int),
    "semb_advect": ([vp, vp, vp, vp, vp, vp], C.c_int),
    "semb_convdiff_create": ([vp, vp, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(vp)], C.c_int),
    "semb_gradT": ([vp, vp, vp, vp], C.c_int),
    "semb_approx_hlmz_inv": ([vp, vp, C.c_double, C.c_char_p, vp], C.c_int),
    "semb_stokes_create": ([vp, vp, C.c_char_p, C.c_char_p, C.c_double, C.POINTER(vp)], C.c_int),
    "semb_stokes_destroy": ([vp], C.c_int),
    "semb_diver": ([vp, vp, vp, vp], C.c_int),
    "semb_diverT": ([vp, vp, vp, vp], C.c_int),
    "semb_stokes_op": ([vp, vp, vp], C.c_int),
    "semb_stokes_rhs": ([vp, vp, vp, vp], C.c_int),
    "semb_stokes_solve": ([vp, vp, vp, C.c_double, C.c_longlong, c_ll_p, c_double_p], C.c_int),
    "semb_stokes_project": ([vp, vp, vp, vp, C.c_double, C.c_longlong, c_ll_p, c_double_p], C.c_int),
    "semb_lapl_host": ([vp, c_double_p, c_double_p], C.c_int),
    "semb_hlmz_host": ([vp, c_double_p, c_double_p, C.c_double, c_double_p, C.c_double, c_double_p], C.c_int),
    "semb_mass_host": ([vp, c_double_p, c_double_p], C.c_int),
    "semb_gather_scatter_host": ([vp, c_double_p, c_double_p], C.c_int),
    "semb_mask_host": ([vp, c_double_p, c_double_p, c_double_p], C.c_int),
    "semb_oplhs_host": ([vp, c_double_p, c_double_p, C.c_double, c_double_p, C.c_double, C.c_char_p, c_double_p,
                         c_double_p], C.c_int),
    "semb_pcg_host": ([vp, C.POINTER(PcgOpts), c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_ll_p,
                       c_double_p], C.c_int),
    "semb_abu_host": ([vp, c_double_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, c_double_p, C.c_int,
                       C.c_int, c_double_p], C.c_int),
    "semb_laplace_host": ([vp, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p,
                           C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p], C.c_int),
    "semb_mass_explicit_host": ([vp, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int,
                                 c_double_p, c_double_p, c_double_p], C.c_int),
    "semb_mul_host": ([vp, C.c_size_t, c_double_p, c_double_p, c_double_p], C.c_int),
    "semb_strip_kernel_info": ([C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p], C.c_int),
    "semb_mesh_plan": ([vp] + [c_int_p] * 5, C.c_int),
    "semb_mesh_set_chunks": ([vp, C.c_int], C.c_int),
    "semb_fdm_create": ([vp, C.c_char_p, C.c_double, C.c_double, C.POINTER(vp)], C.c_int),
    "semb_fdm_destroy": ([vp], C.c_int),
    "semb_fdm_apply": ([vp, vp, vp], C.c_int),
    "semb_fdm_apply_host": ([vp, c_double_p, c_double_p], C.c_int),
    "semb_fdm_tables": ([C.c_int, c_double_p, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p], C.c_int),
    "semb_mesh_fused_tail": ([vp, c_int_p], C.c_int),
    "semb_mesh_groups": ([vp, c_int_p], C.c_int),
    "semb_mesh_debug_read": ([vp, c_ll_p, C.c_int], C.c_int),
    "semb_mesh_peer_status": ([vp], C.c_int),
}

_lib = None


def load():
    """Load libsemb.so (once).  Raises ImportError when it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libsemb.so not found at %s -- build it with `python spectralelements.jl_b200/build.py` "
            "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (argt, rest) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.argtypes = argt
        fn.restype = rest
    _lib = lib
    return lib


def check(rc):
    """Raise SembError for negative codes; pass 0 / +1 (not converged) through."""
    if rc < 0:
        raise SembError(rc, load().semb_last_error().decode("utf-8", "replace"))
    return rc


def dptr(a):
    """double* of a float64 Fortran-contiguous (or 1-D contiguous) NumPy array; None -> NULL."""
    if a is None:
        return None
    assert a.dtype == np.float64
    return a.ctypes.data_as(c_double_p)


def as_f64(a, shape=None):
    """Column-major float64 copy/view of `a` (Bool masks are widened, as the Julia shim must)."""
    b = np.asfortranarray(a, dtype=np.float64)
    if shape is not None and tuple(b.shape) != tuple(shape):
        raise ValueError("DimensionMismatch: expected %s, got %s" % (tuple(shape), tuple(b.shape)))
    return b
