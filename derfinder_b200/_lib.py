"""ctypes binding of libderfinder_b200.so (the C ABI in include/derfinder_b200.h).

The library is the product; this module only marshals numpy arrays to the plain-C entry points.
There is no fallback: if the shared library is missing or no CUDA device exists, the first call
raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libderfinder_b200.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_vpp = C.POINTER(C.c_void_p)


class DfbError(RuntimeError):
    """Raised for every non-zero return code; carries dfb_last_error()."""

    def __init__(self, code, msg):
        super().__init__(f"[derfinder_b200 rc={code}] {msg}")
        self.code = code
        self.msg = msg


DFB_ENOREGIONS = 1
DFB_I32, DFB_F64 = 0, 1
FILTER_NONE, FILTER_ONE, FILTER_MEAN = 0, 1, 2
K_PREP, K_FSTAT, K_FSTAT_TC, K_REGIONS, K_PVAL, K_REGMAT = 1, 2, 3, 4, 5, 6

# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "dfb_last_error": (C.c_char_p, []),
    "dfb_version": (C.c_int, []),
    "dfb_device_count": (C.c_int, []),
    "dfb_create": (C.c_int, [C.c_int, C.c_void_p, c_vpp]),
    "dfb_destroy": (None, [C.c_void_p]),
    "dfb_sync": (C.c_int, [C.c_void_p]),
    "dfb_launch_count": (C.c_int64, [C.c_void_p]),
    "dfb_reset_counters": (None, [C.c_void_p]),
    "dfb_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "dfb_kernel_time": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_i64p]),
    "dfb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "dfb_filter_rle": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_vpp, c_vpp, c_i64p, C.c_int64, c_f64p, C.c_int,
                                 C.c_double, c_i32p, C.c_int64, C.c_int, c_vpp]),
    "dfb_cov_from_dense": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, c_i32p,
                                     C.c_int64, C.c_int, c_vpp]),
    "dfb_cov_from_rle": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_vpp, c_vpp, c_i64p, C.c_int64, c_i32p, C.c_int64,
                                   C.c_int, c_vpp]),
    "dfb_cov_from_dense_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, c_i32p,
                                         C.c_int64, C.c_int, c_vpp]),
    "dfb_cov_view_dense_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, c_i32p,
                                         C.c_int64, C.c_int, c_vpp]),
    "dfb_cov_free": (None, [C.c_void_p]),
    "dfb_cov_info": (C.c_int, [C.c_void_p, c_i64p, C.POINTER(C.c_int), c_i64p, C.POINTER(C.c_int), c_i64p,
                               C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dfb_cov_get_position": (C.c_int, [C.c_void_p, c_i32p, C.POINTER(C.c_int)]),
    "dfb_cov_get_column": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "dfb_cov_get_mean": (C.c_int, [C.c_void_p, c_f64p]),
    "dfb_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, c_i32p, C.c_int]),
    "dfb_cov_get_processed_column": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "dfb_cov_get_group_mean": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "dfb_chunk_lengths": (C.c_int64, [C.c_int64, C.c_double, C.c_int, c_i64p, C.c_int64]),
    "dfb_fstats": (C.c_int, [C.c_void_p, C.c_void_p, c_f64p, C.c_int, c_f64p, C.c_int, C.c_double, c_f64p]),
    "dfb_fstats_begin": (C.c_int, [C.c_void_p, C.c_void_p, c_f64p, C.c_int, c_f64p, C.c_int, C.c_double, c_f64p]),
    "dfb_fstats_wait": (C.c_int, [C.c_void_p]),
    "dfb_model_basis": (C.c_int, [c_f64p, C.c_int, C.c_int, c_f64p, C.c_int, c_f64p, C.POINTER(C.c_int)]),
    "dfb_find_regions": (C.c_int, [C.c_void_p, C.c_void_p, c_f64p, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                   C.c_int, c_vpp]),
    "dfb_find_regions_pos": (C.c_int, [C.c_void_p, c_f64p, C.c_int64, c_i32p, C.c_int64, C.c_int, C.c_double,
                                       C.c_double, C.c_int32, C.c_int32, C.c_int, c_vpp]),
    "dfb_regions_free": (None, [C.c_void_p]),
    "dfb_regions_count": (C.c_int64, [C.c_void_p]),
    "dfb_regions_count_up": (C.c_int64, [C.c_void_p]),
    "dfb_regions_get": (C.c_int, [C.c_void_p, c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, c_i32p, c_i32p, c_f64p]),
    "dfb_regions_view_means": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_f64p]),
    "dfb_regions_view_means_host": (C.c_int, [C.c_void_p, C.c_void_p, c_f64p, C.c_int64, c_f64p]),
    "dfb_cov_from_bigwig": (C.c_int, [C.c_void_p, C.c_int, c_vpp, c_vpp, c_i64p, C.POINTER(C.c_uint32), c_i32p, c_i32p,
                                      C.c_int64, c_f64p, c_vpp]),
    "dfb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "dfb_comm_init_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "dfb_comm_init_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "dfb_comm_destroy": (None, [C.c_void_p]),
    "dfb_comm_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dfb_merge_results": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int64,
                                    C.c_double, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int64)]),
    "dfb_cluster_maker": (C.c_int64, [c_i32p, C.c_int64, C.c_int, C.c_int32, c_i64p, c_i64p, c_i64p, C.c_int64]),
    "dfb_perm_nulls": (C.c_int, [C.c_void_p, C.c_void_p, c_f64p, C.c_int, c_f64p, C.c_int, C.c_double, c_i32p,
                                 C.c_int64, C.c_double, C.c_double, C.c_int32, c_vpp]),
    "dfb_debug_screen": (C.c_int, [C.c_void_p, C.c_void_p, c_f64p, C.c_int, c_f64p, C.c_int, C.c_double, c_i32p,
                                   C.c_int64, C.c_double, c_i64p, c_i64p, c_i64p, c_i64p]),
    "dfb_nulls_free": (None, [C.c_void_p]),
    "dfb_nulls_count": (C.c_int64, [C.c_void_p]),
    "dfb_nulls_counts_per_perm": (C.c_int, [C.c_void_p, c_i64p]),
    "dfb_nulls_get": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_i32p, c_f64p, c_i32p]),
    "dfb_nulls_copy_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfb_pool_local_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p]),
    "dfb_pool_hist_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfb_pool_rank_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "dfb_nulls_merge_areas_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dfb_nulls_sorted_areas_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfb_regions_null_hist_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "dfb_regions_hist_counts_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfb_count_greater_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "dfb_regions_copy_areas_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfb_calc_pval": (C.c_int, [C.c_void_p, c_f64p, C.c_int64, c_f64p, C.c_int64, C.c_int, c_f64p]),
    "dfb_calc_pval_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "dfb_calc_fwer": (C.c_int, [C.c_void_p, C.c_double, c_f64p, c_i32p, C.c_int64, C.c_int64, c_f64p, C.c_int64,
                                c_f64p]),
    "dfb_perm_max_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p]),
    "dfb_quantile": (C.c_int, [C.c_void_p, C.c_void_p, c_f64p, C.c_int64, C.c_double, c_f64p]),
    "dfb_calculate_pvalues": (C.c_int, [C.c_void_p, C.c_void_p, c_f64p, c_f64p, C.c_int, c_f64p, C.c_int, C.c_double,
                                        c_i32p, C.c_int64, C.c_double, C.c_double, C.c_int32, C.c_int32, c_vpp, c_vpp,
                                        c_f64p]),
    "dfb_regions_get_pvalues": (C.c_int, [C.c_void_p, c_f64p]),
    "dfb_regions_get_order": (C.c_int, [C.c_void_p, c_i32p]),
    "dfb_region_matrix": (C.c_int, [C.c_void_p, C.c_void_p, c_i32p, c_i32p, C.c_int64, c_f64p, C.c_int, c_f64p]),
    "dfb_view_sums": (C.c_int, [C.c_void_p, C.c_void_p, c_i32p, c_i32p, C.c_int64, c_f64p, C.c_int, c_f64p]),
    "dfb_region_matrix_r": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_f64p, C.c_int, c_f64p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m derfinder_b200.build` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Raise on negative codes; pass positive informational codes through."""
    if rc < 0:
        raise DfbError(rc, load().dfb_last_error().decode("utf-8", "replace"))
    return rc


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))
