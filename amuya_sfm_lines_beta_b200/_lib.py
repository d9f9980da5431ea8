"""ctypes binding of liblba.so (the C ABI declared in include/lba.h).

The library is built in-tree (``amuya_sfm_lines_beta_b200/liblba.so``) by
``__graft_entry__.build()`` / ``csrc/Makefile``.  There is no fallback: if the
shared object is missing, or no CUDA device is present when a context is
created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblba.so")

LBA_TRACE_MAX = 512
LBA_K_COUNT = 12
LBA_OK = 0
LBA_ERR_CUDA, LBA_ERR_ARG, LBA_ERR_STATE, LBA_ERR_UNSUPPORTED, LBA_ERR_NONFINITE, LBA_ERR_COMM = -1, -2, -3, -4, -5, -6

EXPORTS = [
    "lba_last_error", "lba_version", "lba_default_options", "lba_create", "lba_destroy", "lba_set_stream",
    "lba_synchronize", "lba_wait_stream", "lba_set_problem", "lba_point_wise_indices", "lba_project", "lba_residuals",
    "lba_residuals_jacobian", "lba_normal_equations", "lba_schur_solve", "lba_lm_solve",
    "lba_comm_unique_id", "lba_comm_init", "lba_comm_destroy", "lba_comm_peer_local", "lba_comm_peer_attach", "lba_launch_count", "lba_set_profiling",
    "lba_get_profile", "lba_kernel_class_name", "lba_measure_fp64_peak", "lba_measure_copy_bandwidth",
    "lba_triangulate_lines", "lba_track_table",
    "lba_lines2d_set", "lba_observations_from_tables", "lba_set_problem_lines2d", "lba_triangulate_tracks",
    "lba_residuals12", "lba_project12", "lba_lm_solve12",
]


class LbaOptions(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32), ("pcg_max_iter", C.c_int32),
        ("ftol", C.c_double), ("xtol", C.c_double), ("gtol", C.c_double),
        ("lambda0", C.c_double), ("lambda_min", C.c_double), ("lambda_max", C.c_double),
        ("pcg_tol", C.c_double), ("min_ratio", C.c_double),
        ("verbose", C.c_int32), ("flags", C.c_int32), ("loss", C.c_int32), ("pad", C.c_int32),
        ("f_scale", C.c_double), ("max_evals", C.c_int32), ("max_backtrack", C.c_int32),
    ]


LOSSES = {"linear": 0, "soft_l1": 1, "huber": 2, "cauchy": 3, "arctan": 4}  # LBA_LOSS_* in include/lba.h


class LbaReport(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("status", C.c_int32), ("n_jac", C.c_int32), ("n_res", C.c_int32),
        ("pcg_iters_total", C.c_int32), ("chol_clamped", C.c_int32), ("n_accepted", C.c_int32), ("pad", C.c_int32),
        ("cost_init", C.c_double), ("cost", C.c_double), ("seconds", C.c_double),
        ("trace_len", C.c_int32), ("trace_it", C.c_int32 * LBA_TRACE_MAX),
        ("trace_pcg", C.c_int32 * LBA_TRACE_MAX), ("trace_accepted", C.c_int32 * LBA_TRACE_MAX),
        ("trace_cost", C.c_double * LBA_TRACE_MAX), ("trace_cost_new", C.c_double * LBA_TRACE_MAX),
        ("trace_lambda", C.c_double * LBA_TRACE_MAX), ("trace_pred", C.c_double * LBA_TRACE_MAX),
        ("trace_step_norm", C.c_double * LBA_TRACE_MAX), ("trace_g_norm", C.c_double * LBA_TRACE_MAX),
        ("trace_alpha", C.c_double * LBA_TRACE_MAX),
    ]


class LbaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblba error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """dlopen liblba.so and declare the prototypes.  Loudly fails if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C amuya_sfm_lines_beta_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.lba_last_error.restype = C.c_char_p
    lib.lba_last_error.argtypes = []
    lib.lba_version.restype = C.c_int
    lib.lba_default_options.restype = None
    lib.lba_default_options.argtypes = [C.POINTER(LbaOptions)]
    lib.lba_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.lba_destroy.argtypes = [vp]
    lib.lba_set_stream.argtypes = [vp, vp]
    lib.lba_synchronize.argtypes = [vp]
    lib.lba_wait_stream.argtypes = [vp, vp]
    lib.lba_set_problem.argtypes = [vp, i64, i64, i64, vp, vp, vp, vp, vp]
    lib.lba_point_wise_indices.argtypes = [vp, vp, vp]
    lib.lba_project.argtypes = [vp, vp, vp, vp]
    lib.lba_residuals.argtypes = [vp, vp, vp, vp, C.POINTER(dbl)]
    lib.lba_residuals_jacobian.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.lba_normal_equations.argtypes = [vp, vp, vp, vp, vp]
    lib.lba_schur_solve.argtypes = [vp, dbl, dbl, i32, vp, vp, C.POINTER(i32)]
    lib.lba_lm_solve.argtypes = [vp, vp, vp, C.POINTER(LbaOptions), C.POINTER(LbaReport), vp]
    lib.lba_comm_unique_id.argtypes = [vp]
    lib.lba_comm_init.argtypes = [vp, vp, i32, i32]
    lib.lba_comm_destroy.argtypes = [vp]
    lib.lba_comm_peer_local.argtypes = [vp, vp, i64]
    lib.lba_comm_peer_attach.argtypes = [vp, vp, i32, i32]
    lib.lba_launch_count.restype = i64
    lib.lba_launch_count.argtypes = [vp]
    lib.lba_set_profiling.argtypes = [vp, i32]
    lib.lba_get_profile.argtypes = [vp, C.POINTER(dbl), C.POINTER(i64)]
    lib.lba_kernel_class_name.restype = C.c_char_p
    lib.lba_kernel_class_name.argtypes = [i32]
    lib.lba_measure_fp64_peak.argtypes = [vp, C.POINTER(dbl)]
    lib.lba_measure_copy_bandwidth.argtypes = [vp, i64, C.POINTER(dbl)]
    lib.lba_triangulate_lines.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(dbl)]
    lib.lba_track_table.argtypes = [vp, i32, i64, vp, vp, vp, vp, dbl, vp, i64, vp, vp, vp, C.POINTER(i64), vp]
    lib.lba_lines2d_set.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.lba_observations_from_tables.argtypes = [vp, i64, i32, vp, vp, i64, vp, vp, vp, C.POINTER(i64), C.POINTER(i64)]
    lib.lba_set_problem_lines2d.argtypes = [vp, i64, i64, i64, vp, vp, vp]
    lib.lba_triangulate_tracks.argtypes = [vp, i32, vp, vp, vp, vp, vp, C.POINTER(i64)]
    lib.lba_residuals12.argtypes = [vp, vp, vp, vp, C.POINTER(dbl)]
    lib.lba_project12.argtypes = [vp, i64, vp, vp, vp]
    lib.lba_lm_solve12.argtypes = [vp, vp, vp, C.POINTER(LbaOptions), C.POINTER(LbaReport), vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("lba_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(code):
    if code != LBA_OK:
        raise LbaError(code, load().lba_last_error().decode("utf-8", "replace"))
