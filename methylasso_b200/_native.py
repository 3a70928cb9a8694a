"""ctypes loader for libmethylasso_b200.so (the C ABI of include/methylasso_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is present the import of
the library / creation of an engine raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmethylasso_b200.so")

ABI_VERSION = 1


class MlParams(C.Structure):
    """ml_params (include/methylasso_b200.h)."""
    _fields_ = [("lambda2", C.c_double), ("max_lambda2", C.c_double), ("lambda1", C.c_double),
                ("tol_val", C.c_double), ("bf_per_kb", C.c_double), ("hyper", C.c_double),
                ("max_iter", C.c_uint32), ("nouter", C.c_uint32), ("ref", C.c_uint32),
                ("model", C.c_int32), ("optimize_lambda2", C.c_int32), ("optimize_hyper", C.c_int32)]


class MlLmrParams(C.Structure):
    """ml_lmr_params (include/methylasso_b200.h): thresholds of R/call.R:75-91 with the reference's defaults."""
    _fields_ = [(k, C.c_double) for k in ("min_width", "max_distance", "flank_width", "flank_dist_max", "lmr_max_beta",
                                          "flank_beta_min", "umr_max_beta", "umr_std_threshold", "umr_large_width",
                                          "umr_large_density")] + [("min_num_cpgs", C.c_int32), ("flank_num_cpgs", C.c_int32)]

    def __init__(self, **kw):
        d = dict(min_width=30.0, max_distance=500.0, flank_width=300.0, flank_dist_max=5000.0, lmr_max_beta=0.5,
                 flank_beta_min=0.5, umr_max_beta=0.1, umr_std_threshold=0.1, umr_large_width=500.0,
                 umr_large_density=0.03, min_num_cpgs=4, flank_num_cpgs=10)
        d.update(kw)
        super().__init__(**d)


class MlValleyParams(C.Structure):
    """ml_valley_params (include/methylasso_b200.h): thresholds of R/call.R:100-133 with the reference's defaults."""
    _fields_ = [("valley_max_beta", C.c_double), ("pmd_valley_min_width", C.c_double), ("max_distance", C.c_double),
                ("min_num_cpgs", C.c_int32), ("reserved", C.c_int32)]

    def __init__(self, **kw):
        d = dict(valley_max_beta=0.1, pmd_valley_min_width=5000.0, max_distance=500.0, min_num_cpgs=4, reserved=0)
        d.update(kw)
        super().__init__(**d)


class MlPmdParams(C.Structure):
    """ml_pmd_params (include/methylasso_b200.h): R/call.R:31-35 defaults for the PMD calls."""
    _fields_ = [("pmd_lambda2", C.c_double), ("pmd_std_threshold", C.c_double), ("pmd_max_beta", C.c_double),
                ("split_pmds", C.c_int32), ("merge_pmds", C.c_int32)]

    def __init__(self, **kw):
        d = dict(pmd_lambda2=1000.0, pmd_std_threshold=0.15, pmd_max_beta=0.75, split_pmds=1, merge_pmds=1)
        d.update(kw)
        super().__init__(**d)


# every symbol the header declares, with (restype, argtypes)
_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
SIGNATURES = {
    "ml_abi_version": (C.c_int, []),
    "ml_last_error": (C.c_char_p, []),
    "ml_strerror": (C.c_char_p, [C.c_int]),
    "ml_engine_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ml_engine_destroy": (None, [_P]),
    "ml_engine_set": (C.c_int, [_P, C.c_char_p, C.c_longlong]),
    "ml_engine_launch_count": (C.c_longlong, [_P]),
    "ml_engine_stream": (_P, [_P]),
    "ml_engine_synchronize": (C.c_int, [_P]),
    "ml_engine_set_stream": (C.c_int, [_P, _P]),
    "ml_engine_profile_count": (C.c_int, [_P]),
    "ml_engine_profile_get": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double),
                                        C.POINTER(C.c_longlong)]),
    "ml_engine_profile_work": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "ml_engine_profile_reset": (C.c_int, [_P]),
    "ml_engine_profile_spans": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ml_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(_P)]),
    "ml_host_free": (None, [_P]),
    "ml_weighted_1d_flsa": (C.c_int, [_P, C.c_int64, _P, _P, C.c_double, C.c_double, _P]),
    "ml_weighted_1d_flsa_batch": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_double, _P]),
    "ml_weighted_1d_flsa_batch_device": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_double, _P]),
    "ml_batch_upload": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P), _P]),
    "ml_batch_stage": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "ml_batch_commit": (C.c_int, [_P, _P]),
    "ml_batch_upload_runs": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "ml_batch_free": (None, [_P]),
    "ml_batch_detect": (C.c_int, [_P, _P, C.POINTER(MlParams), C.POINTER(_P), _P]),
    "ml_detect_batch": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.POINTER(MlParams),
                                  C.POINTER(_P), _P]),
    "ml_result_njobs": (C.c_int, [_P]),
    "ml_result_job_sizes": (C.c_int, [_P, C.c_int, _I64P, _I64P, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ml_result_job_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ml_result_copy_job": (C.c_int, [_P, _P, C.c_int] + [_P] * 8),
    "ml_result_copy_all": (C.c_int, [_P, _P] + [_P] * 7),
    "ml_result_copy_all_async": (C.c_int, [_P, _P] + [_P] * 7),
    "ml_result_free": (None, [_P]),
    "ml_result_fast_diff_combine": (C.c_int, [_P, _P]),
    "ml_segment_methylation_helper": (C.c_int, [_P, C.c_int64] + [_P] * 5 + [C.c_double, _I64P] + [_P] * 8),
    "ml_result_segments": (C.c_int, [_P, _P, C.c_double, _I64P] + [_P] * 9),
    "ml_result_pmd_segments": (C.c_int, [_P, _P, C.c_double, C.c_double, _I64P] + [_P] * 9),
    "ml_result_call_table": (C.c_int, [_P, _P, C.c_double, C.c_double, _I64P] + [_P] * 11),
    "ml_result_call_pmd_segments": (C.c_int, [_P, _P, C.c_double, C.c_double, C.c_double, _I64P] + [_P] * 9),
    "ml_wilcoxon_groups": (C.c_int, [_P, C.c_int64] + [_P] * 6),
    "ml_p_adjust_fdr": (C.c_int, [_P, C.c_int64, _P, _P]),
    "ml_result_call_differences": (C.c_int, [_P, _P, _P, C.c_int] + [C.c_double] * 4 + [_I64P] + [_P] * 20),
    "ml_result_call_lmr_umr": (C.c_int, [_P, _P, C.c_double, C.c_double, C.POINTER(MlLmrParams), _I64P] + [_P] * 7),
    "ml_result_call_valleys": (C.c_int, [_P, _P, C.c_double, C.c_double, C.POINTER(MlLmrParams), C.POINTER(MlValleyParams),
                                         _I64P] + [_P] * 9 + [_I64P] + [_P] * 3),
    "ml_result_call_pmd_split": (C.c_int, [_P, _P, C.c_double, C.c_double, C.c_double, C.POINTER(MlLmrParams),
                                           C.POINTER(MlValleyParams), C.c_int, _I64P] + [_P] * 9 + [_I64P] + [_P] * 7),
    "ml_result_call_pmds": (C.c_int, [_P, _P, C.c_double, C.c_double, C.POINTER(MlLmrParams), C.POINTER(MlValleyParams),
                                      C.POINTER(MlPmdParams), _I64P] + [_P] * 11),
    "ml_result_call_pmd_status": (C.c_int, [_P, _P, C.c_double, C.c_double, C.POINTER(MlLmrParams), C.POINTER(MlValleyParams),
                                            C.POINTER(MlPmdParams), _I64P] + [_P] * 9),
    "ml_result_segment_methylation": (C.c_int, [_P, _P, C.c_double, C.c_double, C.POINTER(MlLmrParams),
                                                C.POINTER(MlValleyParams), C.POINTER(MlPmdParams), C.c_int, _I64P] + [_P] * 13
                                      + [_I64P] + [_P] * 11),
    "ml_result_lambda_sweep": (C.c_int, [_P, _P, C.c_int, _P, C.c_double, C.c_double, _P, _P]),
    "ml_gather_layout": (C.c_int, [C.c_int64, C.c_int64, _P, _I64P]),
    "ml_gather_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.POINTER(_P)]),
    "ml_gather_ipc_handle": (C.c_int, [_P, _P]),
    "ml_gather_open": (C.c_int, [_P, _P]),
    "ml_gather_window": (C.c_int, [_P, C.c_int, C.POINTER(_P), _I64P]),
    "ml_gather_push": (C.c_int, [_P, _P, _P, C.c_double, C.c_double, C.POINTER(MlLmrParams), C.POINTER(MlValleyParams),
                                 C.POINTER(MlPmdParams), C.c_int, _P, C.c_int, C.c_int, _I64P, _I64P]),
    "ml_gather_fetch": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ml_gather_destroy": (None, [_P]),
    "ml_text_parse": (C.c_int, [_P, C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_char, C.POINTER(_P), _I64P, _I64P]),
    "ml_text_fetch": (C.c_int, [_P, _P] + [_P] * 8),
    "ml_text_free": (None, [_P]),
    "ml_segment_call_table": (C.c_int, [_P, C.c_int64] + [_P] * 4 + [C.c_int64] + [_P] * 5 + [C.c_double, _I64P] + [_P] * 10),
}

# test hooks of csrc/debug_api.h: exported, but not part of the drop-in surface
DEBUG_SIGNATURES = {
    "ml_debug_flsa_backpointers": (C.c_int, [_P, C.c_int64, _P, _P, C.c_double, _P, _P, _P]),
    "ml_debug_flsa_last_stats": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "ml_debug_div_counts": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "ml_debug_scan_runs": (C.c_int, [C.c_int64, _P, _P, _P, _P, C.c_int, C.c_int64, _I64P, _P, _P, _P,
                                     C.POINTER(C.c_int32), _I64P]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m methylasso_b200.build` "
            "(nvcc, sm_100a). methylasso_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(DEBUG_SIGNATURES.items()):
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ml_abi_version() != ABI_VERSION:
        raise ImportError("libmethylasso_b200.so ABI version mismatch")
    _lib = lib
    return lib
