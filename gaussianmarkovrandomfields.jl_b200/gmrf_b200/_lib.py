"""ctypes binding of libgmrf_b200.so (include/gmrf_b200.h). Fails loudly if the library is missing:
there is no CPU fallback for the numeric path."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libgmrf_b200.so"))

c_i64 = ctypes.c_int64
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p

INFO_KEYS = ["n", "nnz_q", "nnz_l", "nnz_l_stored", "nsuper", "nlevels", "max_front", "max_ns", "update_pool",
             "flops_chol", "flops_chol_stored", "device_bytes", "graph_nodes", "selinv_nodes", "pattern_cache_hits",
             "large_tile_launches", "splitk_tasks", "fast_roots", "chain_launches", "front_launches"]

ORDER_NATURAL, ORDER_ND, ORDER_AMD = 0, 1, 2

EXPORTS = [
    "gmrf_b200_create", "gmrf_b200_destroy", "gmrf_b200_last_error", "gmrf_b200_refactorize",
    "gmrf_b200_refactorize_device", "gmrf_b200_logdet", "gmrf_b200_solve", "gmrf_b200_solve_Lt",
    "gmrf_b200_solve_device", "gmrf_b200_solve_Lt_device", "gmrf_b200_selinv_compute", "gmrf_b200_selinv_diag",
    "gmrf_b200_selinv_nnz", "gmrf_b200_selinv_pattern", "gmrf_b200_selinv_values", "gmrf_b200_selinv_extract",
    "gmrf_b200_info", "gmrf_b200_get_perm", "gmrf_b200_get_colcounts", "gmrf_b200_get_etree",
    "gmrf_b200_get_supernodes", "gmrf_b200_get_rows", "gmrf_b200_get_scatter", "gmrf_b200_last_timings",
    "gmrf_b200_get_factor_panels", "gmrf_b200_get_selinv_panels", "gmrf_b200_set_option",
    "gmrf_b200_test_gemm", "gmrf_b200_test_potrf", "gmrf_b200_test_potrf_inv", "gmrf_b200_bench_gemm", "gmrf_b200_profile_refactorize", "gmrf_b200_profile_plan", "gmrf_b200_device_array", "gmrf_b200_set_value_basis", "gmrf_b200_set_base_values", "gmrf_b200_refactorize_base_minus_diag", "gmrf_b200_lane_capacity", "gmrf_b200_refactorize_lanes", "gmrf_b200_refactorize_combination_lanes", "gmrf_b200_refactorize_combination", "gmrf_b200_adopt_factor", "gmrf_b200_host_register", "gmrf_b200_host_unregister",
    "gmrf_b200_selinv_dot", "gmrf_b200_selinv_dot_basis",
    "gmrf_b200_factor_nnz", "gmrf_b200_factor_pattern", "gmrf_b200_factor_values", "gmrf_b200_pattern_positions",
    "gmrf_b200_create_from_analysis", "gmrf_b200_analysis_export", "gmrf_b200_analysis_equal",
    "gmrf_b200_test_rsqrt", "gmrf_b200_host_copy", "gmrf_b200_debug_chain_phases", "gmrf_b200_analysis_fingerprint", "gmrf_b200_adopt_factor_checked",
    "gmrf_b200_plan_launch_info", "gmrf_b200_set_hessian_pattern", "gmrf_b200_refactorize_base_minus_sparse", "gmrf_b200_selinv_quadform_rows",
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(gmrf_b200 has no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    L.gmrf_b200_create.restype = ctypes.c_int
    L.gmrf_b200_create.argtypes = [ctypes.POINTER(c_vp), c_i64, c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int]
    L.gmrf_b200_destroy.restype = None
    L.gmrf_b200_destroy.argtypes = [c_vp]
    L.gmrf_b200_last_error.restype = ctypes.c_char_p
    L.gmrf_b200_last_error.argtypes = [c_vp]
    for name in ("gmrf_b200_refactorize", "gmrf_b200_refactorize_device"):
        f = getattr(L, name)
        f.restype = ctypes.c_int
        f.argtypes = [c_vp, c_vp, c_i64]
    L.gmrf_b200_logdet.restype = ctypes.c_int
    L.gmrf_b200_logdet.argtypes = [c_vp, c_f64p]
    for name in ("gmrf_b200_solve", "gmrf_b200_solve_Lt", "gmrf_b200_solve_device", "gmrf_b200_solve_Lt_device"):
        f = getattr(L, name)
        f.restype = ctypes.c_int
        f.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64]
    L.gmrf_b200_selinv_compute.restype = ctypes.c_int
    L.gmrf_b200_selinv_compute.argtypes = [c_vp]
    L.gmrf_b200_selinv_diag.restype = ctypes.c_int
    L.gmrf_b200_selinv_diag.argtypes = [c_vp, c_vp]
    L.gmrf_b200_selinv_nnz.restype = ctypes.c_int
    L.gmrf_b200_selinv_nnz.argtypes = [c_vp, c_i64p]
    L.gmrf_b200_selinv_pattern.restype = ctypes.c_int
    L.gmrf_b200_selinv_pattern.argtypes = [c_vp, c_vp, c_vp, ctypes.c_int]
    L.gmrf_b200_selinv_values.restype = ctypes.c_int
    L.gmrf_b200_selinv_values.argtypes = [c_vp, c_vp]
    L.gmrf_b200_selinv_extract.restype = ctypes.c_int
    L.gmrf_b200_selinv_extract.argtypes = [c_vp, c_i64, c_vp, c_vp, ctypes.c_int, c_vp]
    L.gmrf_b200_create_from_analysis.restype = ctypes.c_int
    L.gmrf_b200_create_from_analysis.argtypes = [ctypes.POINTER(c_vp), c_i64, c_vp, c_vp, ctypes.c_int, c_vp, c_i64, ctypes.c_int]
    L.gmrf_b200_analysis_export.restype = ctypes.c_int
    L.gmrf_b200_analysis_export.argtypes = [c_vp, c_vp, c_i64, c_i64p]
    L.gmrf_b200_analysis_equal.restype = ctypes.c_int
    L.gmrf_b200_analysis_equal.argtypes = [c_vp, c_vp]
    L.gmrf_b200_pattern_positions.restype = ctypes.c_int
    L.gmrf_b200_pattern_positions.argtypes = [c_vp, c_i64, c_vp, c_vp, ctypes.c_int, c_vp]
    L.gmrf_b200_factor_nnz.restype = ctypes.c_int
    L.gmrf_b200_factor_nnz.argtypes = [c_vp, c_i64p]
    L.gmrf_b200_factor_pattern.restype = ctypes.c_int
    L.gmrf_b200_factor_pattern.argtypes = [c_vp, c_vp, c_vp, ctypes.c_int]
    L.gmrf_b200_factor_values.restype = ctypes.c_int
    L.gmrf_b200_factor_values.argtypes = [c_vp, c_vp]
    L.gmrf_b200_selinv_dot.restype = ctypes.c_int
    L.gmrf_b200_selinv_dot.argtypes = [c_vp, c_i64, c_vp, c_vp, ctypes.c_int, c_vp, c_f64p]
    L.gmrf_b200_selinv_dot_basis.restype = ctypes.c_int
    L.gmrf_b200_selinv_dot_basis.argtypes = [c_vp, c_vp, ctypes.c_int]
    L.gmrf_b200_info.restype = ctypes.c_int
    L.gmrf_b200_info.argtypes = [c_vp, c_vp, ctypes.c_int]
    L.gmrf_b200_get_perm.restype = ctypes.c_int
    L.gmrf_b200_get_perm.argtypes = [c_vp, c_vp, ctypes.c_int]
    for name in ("gmrf_b200_get_colcounts", "gmrf_b200_get_etree"):
        f = getattr(L, name)
        f.restype = ctypes.c_int
        f.argtypes = [c_vp, c_vp]
    L.gmrf_b200_get_supernodes.restype = ctypes.c_int
    L.gmrf_b200_get_supernodes.argtypes = [c_vp] + [c_vp] * 8
    L.gmrf_b200_get_rows.restype = ctypes.c_int
    L.gmrf_b200_get_rows.argtypes = [c_vp, c_vp, c_vp]
    L.gmrf_b200_get_scatter.restype = ctypes.c_int
    L.gmrf_b200_get_scatter.argtypes = [c_vp, c_i64p, c_vp, c_vp]
    L.gmrf_b200_last_timings.restype = ctypes.c_int
    L.gmrf_b200_last_timings.argtypes = [c_vp, c_vp, ctypes.c_int]
    for name in ("gmrf_b200_get_factor_panels", "gmrf_b200_get_selinv_panels"):
        f = getattr(L, name)
        f.restype = ctypes.c_int
        f.argtypes = [c_vp, c_vp, c_i64]
    L.gmrf_b200_set_option.restype = ctypes.c_int
    L.gmrf_b200_set_option.argtypes = [ctypes.c_char_p, ctypes.c_double]
    L.gmrf_b200_test_gemm.restype = ctypes.c_int
    L.gmrf_b200_test_gemm.argtypes = [ctypes.c_int] * 7 + [c_vp, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_double, c_vp, ctypes.c_int]
    L.gmrf_b200_test_potrf.restype = ctypes.c_int
    L.gmrf_b200_test_potrf.argtypes = [ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.gmrf_b200_test_potrf_inv.restype = ctypes.c_int
    L.gmrf_b200_test_potrf_inv.argtypes = [ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int, c_vp, ctypes.POINTER(ctypes.c_int)]
    L.gmrf_b200_bench_gemm.restype = ctypes.c_int
    L.gmrf_b200_bench_gemm.argtypes = [ctypes.c_int] * 8 + [c_f64p]
    L.gmrf_b200_profile_refactorize.restype = ctypes.c_int
    L.gmrf_b200_profile_refactorize.argtypes = [c_vp, c_vp, c_vp, c_f64p]
    L.gmrf_b200_profile_plan.restype = ctypes.c_int
    L.gmrf_b200_profile_plan.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, c_i64, c_vp, c_vp, c_vp, c_vp]
    L.gmrf_b200_set_value_basis.restype = ctypes.c_int
    L.gmrf_b200_set_value_basis.argtypes = [c_vp, c_vp, ctypes.c_int]
    L.gmrf_b200_refactorize_combination.restype = ctypes.c_int
    L.gmrf_b200_refactorize_combination.argtypes = [c_vp, c_vp, ctypes.c_int]
    L.gmrf_b200_set_base_values.restype = ctypes.c_int
    L.gmrf_b200_set_base_values.argtypes = [c_vp, c_vp, c_i64]
    L.gmrf_b200_refactorize_base_minus_diag.restype = ctypes.c_int
    L.gmrf_b200_refactorize_base_minus_diag.argtypes = [c_vp, c_vp, c_i64]
    L.gmrf_b200_lane_capacity.restype = ctypes.c_int
    L.gmrf_b200_lane_capacity.argtypes = [c_vp]
    L.gmrf_b200_refactorize_lanes.restype = ctypes.c_int
    L.gmrf_b200_refactorize_lanes.argtypes = [c_vp, c_vp, c_i64, ctypes.c_int, c_vp, c_vp]
    L.gmrf_b200_refactorize_combination_lanes.restype = ctypes.c_int
    L.gmrf_b200_refactorize_combination_lanes.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp]
    L.gmrf_b200_device_array.restype = ctypes.c_int
    L.gmrf_b200_device_array.argtypes = [c_vp, ctypes.c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64)]
    L.gmrf_b200_adopt_factor.restype = ctypes.c_int
    L.gmrf_b200_adopt_factor.argtypes = [c_vp, ctypes.c_double, ctypes.c_int]
    L.gmrf_b200_host_register.restype = ctypes.c_int
    L.gmrf_b200_host_register.argtypes = [c_vp, c_i64]
    L.gmrf_b200_host_copy.restype = ctypes.c_int
    L.gmrf_b200_host_copy.argtypes = [c_vp, c_vp, c_i64]
    L.gmrf_b200_test_rsqrt.restype = ctypes.c_int
    L.gmrf_b200_test_rsqrt.argtypes = [ctypes.c_int, ctypes.c_int, c_vp, c_vp]
    L.gmrf_b200_debug_chain_phases.restype = ctypes.c_int
    L.gmrf_b200_debug_chain_phases.argtypes = [c_vp, c_vp, ctypes.c_int]
    L.gmrf_b200_analysis_fingerprint.restype = ctypes.c_int
    L.gmrf_b200_analysis_fingerprint.argtypes = [c_vp, ctypes.POINTER(ctypes.c_uint64)]
    L.gmrf_b200_adopt_factor_checked.restype = ctypes.c_int
    L.gmrf_b200_adopt_factor_checked.argtypes = [c_vp, ctypes.c_uint64, ctypes.c_double, ctypes.c_int, ctypes.c_int]
    L.gmrf_b200_plan_launch_info.restype = ctypes.c_int
    L.gmrf_b200_plan_launch_info.argtypes = [c_vp, ctypes.c_int, c_i64, c_vp, c_vp, c_vp, c_vp]
    L.gmrf_b200_set_hessian_pattern.restype = ctypes.c_int
    L.gmrf_b200_set_hessian_pattern.argtypes = [c_vp, c_vp, c_i64, ctypes.c_int]
    L.gmrf_b200_refactorize_base_minus_sparse.restype = ctypes.c_int
    L.gmrf_b200_refactorize_base_minus_sparse.argtypes = [c_vp, c_vp, c_i64]
    L.gmrf_b200_selinv_quadform_rows.restype = ctypes.c_int
    L.gmrf_b200_selinv_quadform_rows.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, ctypes.c_int, c_vp]
    L.gmrf_b200_host_unregister.restype = ctypes.c_int
    L.gmrf_b200_host_unregister.argtypes = [c_vp]
    _lib = L
    return L


def ptr(a):
    """Raw pointer of a numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(c_vp)


def set_option(key: str, value: float):
    rc = lib().gmrf_b200_set_option(key.encode(), float(value))
    if rc != 0:
        raise ValueError(f"unknown option {key!r}")


def host_copy(dst, src):
    """dst[:] = src for float64 arrays of equal size; large contiguous ones go through the library's threaded copy (the
    host mirror `ws.Q.nzval .= nzval` of update_precision_values is 520 MB at 1 M dofs)."""
    import numpy as np
    if (dst.size >= (1 << 20) and isinstance(src, np.ndarray) and src.dtype == np.float64 and dst.dtype == np.float64
            and src.size == dst.size and src.flags.c_contiguous and dst.flags.c_contiguous and not np.shares_memory(dst, src)):
        rc = lib().gmrf_b200_host_copy(ptr(dst), ptr(src), dst.size)
        if rc == 0:
            return
    dst[:] = src
