"""ctypes binding of libkexp.so (include/kexp.h).  There is no CPU fallback: if the CUDA
library is missing or no B200 is visible, importing/creating a context fails loudly."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KEXP_LIB", os.path.join(_HERE, "libkexp.so"))     # KEXP_LIB: kernel-variant experiments


class KexpError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_solve", C.c_int64), ("cg_solves", C.c_int64),
                ("cg_iterations", C.c_int64), ("op_applications", C.c_int64),
                ("kernel_launches", C.c_int64), ("bicg_iterations", C.c_int64),
                ("last_max_rel_residual", C.c_double), ("model_bytes", C.c_double),
                ("spmm_cg_bytes", C.c_double), ("spmm_cg_launches", C.c_int64),
                ("unconverged_solves", C.c_int64), ("worst_rel_residual", C.c_double),
                ("mass_iterations_per_solve", C.c_int64), ("h2d_vector_bytes", C.c_int64), ("d2h_vector_bytes", C.c_int64),
                ("solve_path", C.c_int64)]


_p = C.c_void_p
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

# name -> (restype, argtypes); raw pointers are c_void_p so None (NULL) is accepted
PROTOTYPES = {
    "kexp_create": (C.c_int, [C.POINTER(_p), C.c_int]),
    "kexp_destroy": (C.c_int, [_p]),
    "kexp_last_error": (C.c_char_p, [_p]),
    "kexp_version": (C.c_char_p, []),
    "kexp_set_operator": (C.c_int, [_p, C.c_int64, C.c_int64, _p, _p, _p, _p, _p, _p, _p, C.c_int64]),
    "kexp_set_operator_plan": (C.c_int, [_p, C.c_int64, C.c_int64, C.c_int64] + [_p] * 12),
    "kexp_set_mass_solver": (C.c_int, [_p, C.c_double, C.c_int, C.c_int]),
    "kexp_set_mass_solver_chebyshev": (C.c_int, [_p, C.c_double, C.c_double, C.c_double, C.c_int]),
    "kexp_set_cheby": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_double]),
    "kexp_set_cheby_time_range": (C.c_int, [_p, C.c_int, C.c_double]),
    "kexp_series_reexpand": (C.c_int, [_p, C.c_int, C.c_double, C.c_double, _p, C.POINTER(C.c_int)]),
    "kexp_cheby_apply": (C.c_int, [_p, C.c_int, _p, _p, C.c_double]),
    "kexp_cheby_apply_batch": (C.c_int, [_p, C.c_int, _p, C.c_int64, C.c_int, _p, C.c_int64, _p, C.c_int]),
    "kexp_cheby_apply_average": (C.c_int, [_p, C.c_int, _p, C.c_int64, C.c_int, _p, _p, C.c_int, _p, C.c_int]),
    "kexp_cheby_apply_average_fused": (C.c_int, [_p, C.c_int, _p, C.c_int64, C.c_int, _p, _p, C.c_int, _p, C.c_int]),
    "kexp_batch_create": (C.c_int, [_p, C.c_int, C.POINTER(_p)]),
    "kexp_batch_destroy": (C.c_int, [_p]),
    "kexp_batch_num_panels": (C.c_int, [_p]),
    "kexp_batch_set_order": (C.c_int, [_p, _p]),
    "kexp_batch_upload": (C.c_int, [_p, _p, C.c_int64, C.c_int]),
    "kexp_batch_download": (C.c_int, [_p, _p, C.c_int64]),
    "kexp_batch_upload_dev": (C.c_int, [_p, _p, C.c_int64, C.c_int]),
    "kexp_batch_export_dev": (C.c_int, [_p, _p, C.c_int64]),
    "kexp_get_stream": (C.c_int, [_p, C.POINTER(_p)]),
    "kexp_batch_download_async": (C.c_int, [_p, _p, C.c_int64]),
    "kexp_cheby_apply_batch_async": (C.c_int, [_p, C.c_int, _p, C.c_int64, C.c_int, _p, C.c_int64, _p, C.c_int]),
    "kexp_wait_transfers": (C.c_int, [_p]),
    "kexp_cheby_apply_dev": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "kexp_batch_average_dev": (C.c_int, [_p, _p, _p]),
    "kexp_cheby_average_fused_dev": (C.c_int, [_p, C.c_int, _p, _p, _p, _p]),
    "kexp_synchronize": (C.c_int, [_p]),
    "kexp_set_async": (C.c_int, [_p, C.c_int]),
    "kexp_timer_start": (C.c_int, [_p]),
    "kexp_timer_stop": (C.c_int, [_p, C.POINTER(C.c_double)]),
    "kexp_last_device_ms": (C.c_int, [_p, C.POINTER(C.c_double)]),
    "kexp_set_kernel_timing": (C.c_int, [_p, C.c_int]),
    "kexp_get_spmm_cg_ms": (C.c_int, [_p, C.POINTER(C.c_double)]),
    "kexp_operator_apply_dev": (C.c_int, [_p, _p, _p]),
    "kexp_spmm_K_dev": (C.c_int, [_p, _p, _p]),
    "kexp_set_krylov_operator": (C.c_int, [_p, C.c_double] + [_p] * 12 + [C.c_double, C.c_int]),
    "kexp_krylov_apply": (C.c_int, [_p, _p, _p, C.c_double, C.c_int, C.c_int, _p, _p, _p]),
    "kexp_krylov_apply_batch": (C.c_int, [_p, _p, C.c_int64, C.c_int, _p, C.c_int64, _p, C.c_int, C.c_int, C.c_int]),
    "kexp_dense_expm": (C.c_int, [_p, _p, C.c_int, _p]),
    "kexp_slow_form_dev": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [_p] * 15 + [_p, C.c_int64, _p, C.c_int64, C.c_int, C.c_double]),
    "kexp_comm_unique_id": (C.c_int, [_p]),
    "kexp_comm_init": (C.c_int, [_p, _p, C.c_int, C.c_int]),
    "kexp_allreduce_sum_dev": (C.c_int, [_p, _p, C.c_int64]),
    "kexp_stats": (C.c_int, [_p, C.POINTER(Stats)]),
    "kexp_stats_reset": (C.c_int, [_p]),
    "kexp_get_class_stats": (C.c_int, [_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "kexp_get_class_matrix_bytes": (C.c_int, [_p, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load libkexp.so (built by krylov_exp_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KexpError(f"{LIB_PATH} is missing: run `python -m krylov_exp_b200.build` "
                        "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """Raw pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def check(lib, ctx, rc, what):
    if rc != 0:
        msg = lib.kexp_last_error(ctx)
        raise KexpError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
