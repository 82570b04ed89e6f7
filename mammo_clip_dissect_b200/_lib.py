"""ctypes binding of include/mcd_b200.h (libmcd_b200.so, built in-tree by build.py).

There is deliberately no fallback: if the library is missing or a call fails, the caller gets an
exception -- the scoring path never silently runs on the CPU or through stock torch ops.
"""
import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmcd_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "mcd_b200.h")

OK = 0
F32, F16, BF16 = 0, 1, 2
POOL_MEAN, POOL_MAX = 0, 1
LSE_BLOCK = 256

_c = ctypes
_p, _i64, _i32, _f32, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); must list every function declared in include/mcd_b200.h
SIGNATURES = {
    "mcd_abi_version": (_i32, []),
    "mcd_strerror": (_c.c_char_p, [_i32]),
    "mcd_build_info": (_c.c_char_p, []),
    "mcd_launch_count": (_c.c_uint64, []),
    "mcd_device_check": (_i32, []),
    "mcd_set_tunable": (_i32, [_c.c_char_p, _i64]),
    "mcd_softmax_rows_f32": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _f32, _p]),
    "mcd_last_gemm_path": (_i32, []),
    "mcd_gemm_nt_softmax_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "mcd_gemm_nt_softmax_f32": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _i64, _i32, _f32, _p, _i64, _p, _i64, _p, _sz, _p]),
    "mcd_topk_cols_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "mcd_topk_cols_f32": (_i32, [_p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "mcd_wpmi_accum_f32": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _i64, _p, _f32, _p, _i64, _p]),
    "mcd_wpmi_accum_prob_f32": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _i64, _p, _f32, _p, _i64, _p]),
    "mcd_col_lse_partials_f32": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "mcd_pmi_finalize_f32": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _i64, _f32, _p, _p, _i64, _p]),
    "mcd_bcast_f32": (_i32, [_p, _i64, _p, _i32, _i64, _p]),
    "mcd_pmi_finalize_bcast_f32": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _f32, _p, _p, _i32, _i64, _p]),
    "mcd_pmi_scores_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "mcd_pmi_scores_f32": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _f32, _f32, _p, _f32, _p, _i64, _p, _sz, _p]),
    "mcd_pmi_logsums_f32": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _f32, _p, _f32, _p, _i64, _p, _p, _sz, _p]),
    "mcd_col_lse_partials_seg_f32": (_i32, [_p, _i64, _i64, _p, _i64, _p, _p]),
    "mcd_pmi_finalize_seg_f32": (_i32, [_p, _i64, _i64, _p, _p, _i64, _p, _p, _i64, _f32, _p, _p, _i64, _p]),
    "mcd_pmi_finalize_topk_f32": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _i64, _f32, _p, _p, _i64, _i64, _p, _p, _p]),
    "mcd_pmi_finalize_seg_topk_f32": (_i32, [_p, _i64, _i64, _i64, _p, _p, _i64, _p, _p, _i64, _f32, _p, _p, _i64, _i64, _p, _p, _p]),
    "mcd_row_topk_f32": (_i32, [_p, _i64, _i64, _i64, _i64, _p, _p, _p]),
    "mcd_pool_nchw_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "mcd_pool_nchw": (_i32, [_p, _i32, _i64, _i64, _i64, _i64, _i32, _p, _p, _sz, _p]),
    "mcd_pool_nchw_to": (_i32, [_p, _i32, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i32, _i64, _p, _sz, _p]),
    "mcd_col_stats_f32": (_i32, [_p, _i64, _i64, _i64, _i32, _f32, _p, _p, _p]),
    "mcd_rank_reorder_f32": (_i32, [_p, _i64, _i64, _i64, _p, _p, _i64, _i64, _p, _f32, _f32, _p, _sz, _p, _i64, _p]),
    "mcd_rank_baseline_draws_f32": (_i32, [_p, _i64, _i64, _i64, _p, _f32, _p, _sz, _p]),
    "mcd_rank_baseline_perms_f32": (_i32, [_p, _i64, _i64, _i64, _p, _f32, _p, _sz, _p]),
    "mcd_rank_errors_f32": (_i32, [_p, _i64, _i64, _i64, _p, _p, _i64, _i64, _f32, _f32, _p, _sz, _p, _i64, _p]),
    "mcd_rank_finish_f32": (_i32, [_i64, _i64, _i64, _p, _sz, _p, _i64, _p]),
    "mcd_cos_similarity_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "mcd_cos_similarity_f32": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _i64, _i32, _f32, _p, _i64, _p, _sz, _p]),
    "mcd_last_cos_path": (_i32, []),
    "mcd_rank_reorder_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "mcd_mt19937_draws": (_i32, [_p, _i64, _p, _p]),
    "mcd_rank_reorder_draws_f32": (_i32, [_p, _i64, _i64, _i64, _p, _p, _i64, _i64, _p, _f32, _f32, _p, _sz, _p, _i64, _p]),
    "mcd_cos_matmul_f32": (_i32, [_p, _i64, _p, _p, _p, _i64, _p, _p, _i64, _i64, _i64, _i32, _p, _i64, _p]),
}

_lib = None


def declared_symbols():
    """Function names declared in include/mcd_b200.h (used by the CPU test-suite)."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mcd_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "mammo_clip_dissect_b200: %s is missing. Build it with `python -m mammo_clip_dissect_b200.build` "
                "(nvcc, sm_100a). There is no CPU or stock-torch fallback for the scoring path." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.mcd_abi_version() != 1:
            raise ImportError("libmcd_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(code, what):
    if code != OK:
        raise RuntimeError("%s failed: %s (code %d)" % (what, lib().mcd_strerror(code).decode(), code))


def launch_count():
    return int(lib().mcd_launch_count())


def set_tunable(name, value):
    check(lib().mcd_set_tunable(name.encode(), int(value)), "mcd_set_tunable(%s)" % name)
