"""ctypes binding of libfzb200.so (the C ABI declared in include/frankenz_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is usable, the
first compute call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FZB_LIB_PATH") or os.path.join(HERE, "lib", "libfzb200.so")   # override: kernel experiments

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)

PREC_AUTO, PREC_FP64, PREC_FP32 = 0, 1, 2


class FzbConfig(C.Structure):
    _fields_ = [("free_scale", C.c_int32), ("ignore_model_err", C.c_int32), ("dim_prior", C.c_int32),
                ("track_scale", C.c_int32), ("ltol", C.c_double), ("use_wt_thresh", C.c_int32),
                ("use_cdf_thresh", C.c_int32), ("wt_thresh", C.c_double), ("cdf_thresh", C.c_double),
                ("precision", C.c_int32), ("reserved", C.c_int32)]


class FzbFitOut(C.Structure):
    _fields_ = [("lnprior", c_double_p), ("lnlike", c_double_p), ("lnprob", c_double_p), ("Ndim", c_int64_p),
                ("chi2", c_double_p), ("scale", c_double_p), ("scale_err", c_double_p)]


class FzbStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("pairs_fp32", C.c_int64), ("pairs_fp64", C.c_int64),
                ("objects_fp64", C.c_int64), ("ms_scan", C.c_double), ("ms_accum", C.c_double),
                ("ms_finish", C.c_double), ("ms_total", C.c_double), ("sweep_kind", C.c_int64),
                ("knn_redo", C.c_int64), ("pairs_pass2", C.c_int64), ("ms_summarize", C.c_double),
                ("knn_tc", C.c_int64), ("cut_recorded", C.c_int64), ("cut_changed", C.c_int64),
                ("knn_tc_err", C.c_double), ("objects_fused", C.c_int64),
                ("knn_overflow", C.c_int64)]


# name -> (restype, argtypes); every symbol declared in include/frankenz_b200.h
_H = C.c_void_p
_CFG = C.POINTER(FzbConfig)
_OUT = C.POINTER(FzbFitOut)
SIGNATURES = {
    "fzb_last_error": (C.c_char_p, []),
    "fzb_version": (C.c_int, []),
    "fzb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "fzb_create": (C.c_int, [C.c_int, C.POINTER(_H)]),
    "fzb_destroy": (C.c_int, [_H]),
    "fzb_synchronize": (C.c_int, [_H]),
    "fzb_get_stats": (C.c_int, [_H, C.POINTER(FzbStats)]),
    "fzb_measure_peaks": (C.c_int, [_H, C.c_int, c_double_p, c_double_p]),
    "fzb_set_models": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int32]),
    "fzb_set_lnprior": (C.c_int, [_H, c_double_p, C.c_int64]),
    "fzb_set_lnprior_table": (C.c_int, [_H, c_double_p, C.c_int32, C.c_int64]),
    "fzb_set_object_prior_bins": (C.c_int, [_H, c_int32_p, C.c_int64]),
    "fzb_set_kde_dict": (C.c_int, [_H, C.c_int32, C.c_int32, c_int32_p, c_int64_p, c_double_p, c_double_p]),
    "fzb_set_labels_dict": (C.c_int, [_H, c_int64_p, c_int64_p, C.c_int64]),
    "fzb_set_kde_grid": (C.c_int, [_H, c_double_p, C.c_int32]),
    "fzb_set_labels_grid": (C.c_int, [_H, c_double_p, c_double_p, c_int64_p, c_int64_p, C.c_int64]),
    "fzb_fit": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, C.c_int64, _CFG, _OUT]),
    "fzb_fit_predict": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, C.c_int64, _CFG, c_double_p, c_double_p,
                                  c_double_p, c_int64_p, c_double_p, c_double_p]),
    "fzb_fit_predict_summarize": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, C.c_int64, _CFG, c_double_p,
                                            c_double_p, c_double_p, C.c_int32, C.c_double, c_double_p, c_double_p,
                                            c_double_p, c_int64_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                            c_double_p, c_double_p, c_double_p, c_double_p]),
    "fzb_fit_predict_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _CFG, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fzb_predict_logwt": (C.c_int, [_H, c_double_p, C.c_int64, C.c_int64, c_int64_p, c_int64_p, _CFG, c_double_p,
                                    c_double_p, c_double_p]),
    "fzb_shard_pass1_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _CFG, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "fzb_shard_pass2_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _CFG, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "fzb_shard_pass1_packed_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _CFG, C.c_int64,
                                             C.c_void_p]),
    "fzb_shard_merge_dev": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fzb_shard_pass2_f32_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _CFG, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "fzb_shard_normalise_dev": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "fzb_get_stream": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "fzb_knn_build": (C.c_int, [_H, c_float_p, C.c_int32, C.c_int64, C.c_int32]),
    "fzb_knn_query": (C.c_int, [_H, c_double_p, C.c_int64, C.c_int32, C.c_double, c_int64_p, c_double_p]),
    "fzb_knn_fit": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int32, C.c_double,
                              _CFG, c_int64_p, c_int64_p, _OUT]),
    "fzb_knn_fit_predict": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int32, C.c_double,
                                      _CFG, c_double_p, c_double_p, c_double_p, c_int64_p]),
    "fzb_fit_gather": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int64, c_int64_p, c_int64_p, _CFG,
                                 _OUT]),
    "fzb_pdfs_summarize": (C.c_int, [_H, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int32, C.c_int32,
                                     c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "fzb_pdfs_conf": (C.c_int, [_H, c_double_p, c_double_p, C.c_int64, c_double_p]),
    "fzb_nz_set_pdfs": (C.c_int, [_H, c_double_p, C.c_int64, C.c_int32]),
    "fzb_nz_set_pdfs_dev": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int32]),
    "fzb_nz_loglike": (C.c_int, [_H, c_double_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, c_double_p, c_double_p]),
    "fzb_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "fzb_free_pinned": (C.c_int, [C.c_void_p]),
    "fzb_clean_inplace_f64": (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int64]),
}

_lib = None


class FzbError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("frankenz_b200: %s not found. Build it with `python -m frankenz_b200.build` "
                          "(needs nvcc); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().fzb_last_error()
        raise FzbError(msg.decode("utf-8", "replace") if msg else "libfzb200 error %d" % rc)


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int64_p)


def f64(a):
    """C-contiguous float64 copy/view."""
    return np.ascontiguousarray(a, dtype=np.float64)
