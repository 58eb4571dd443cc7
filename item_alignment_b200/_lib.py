"""ctypes binding of libia_b200.so (C ABI: include/ia_b200.h).

This is the whole FFI surface: plain pointers and sizes, no torch types cross it.  There is NO
fallback: if the CUDA library is missing or an entry point fails, the product raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libia_b200.so")
LIB_PATH = os.environ.get("IA_B200_LIB", LIB_PATH)      # A/B builds of the same ABI (scripts/), never a fallback

# enums of include/ia_b200.h
MEASURES = {"inner_product": 0, "cosine": 1, "l1": 2, "l2": 3}
LOSSES = {"bce": 0, "hinge": 1, "euclidean": 2, "cosine": 3}
REDUCTIONS = {"none": 0, "mean": 1, "sum": 2}
IA_F32, IA_BF16, IA_F16, IA_F64 = 0, 1, 2, 3
IA_MAX_K = 128

_lib = None

# name -> (restype, argtypes); every symbol include/ia_b200.h declares
SIGNATURES = {
    "ia_version": (c_char_p, []),
    "ia_last_error": (c_char_p, []),
    "ia_workspace_bytes": (c_size_t, []),
    "ia_launch_count": (c_int64, []),
    "ia_pair_score_fwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                  c_void_p, c_double, c_void_p, c_void_p]),
    "ia_pair_score_loss_fwd_bwd": (c_int, [c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_int64,
                                           c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int64, c_int64, c_float, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "ia_pair_score_loss_act_bwd": (c_int, [c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_int64,
                                           c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int64, c_int64, c_float, c_void_p, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "ia_dropout_fwd": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int64, c_float, c_uint64, c_uint32, c_uint32, c_void_p, c_int64, c_void_p]),
    "ia_project_tanh_dropout_fwd": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                            c_int64, c_void_p, c_void_p, c_int64, c_int64, c_float, c_uint64, c_uint32, c_void_p]),
    "ia_tanh_dropout_bwd": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_int64, c_void_p]),
    "ia_transpose16": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "ia_project_dgrad": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                 c_void_p, c_int64, c_int64, c_float, c_uint64, c_uint32, c_void_p]),
    "ia_project_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "ia_project_wgrad": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "ia_pair_score_bwd": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                  c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "ia_pair_score_gather_fwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                         c_int64, c_int64, c_void_p, c_void_p, c_double, c_void_p, c_void_p]),
    "ia_pair_score_gather_loss_fwd_bwd": (c_int, [c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_int64,
                                                  c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_size_t, c_void_p]),
    "ia_threshold_sweep": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "ia_best_f1_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "ia_best_f1_threshold": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ia_score_loss_fwd_bwd": (c_int, [c_int, c_float, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_float,
                                      c_void_p, c_size_t, c_void_p]),
    "ia_softmax_head_workspace_bytes": (c_size_t, [c_int64]),
    "ia_softmax_head_fwd_bwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                        c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                        c_int64, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "ia_softmax_head_last_stats": (c_int, [c_void_p]),
    "ia_scale_inplace": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "ia_project_tanh_fwd": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                    c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "ia_project_last_stats": (c_int, [c_void_p]),
    "ia_project_score_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "ia_project_score_fwd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                     c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_double,
                                     c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "ia_row_inv_norm": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p]),
    "ia_catalog_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "ia_catalog_destroy": (None, [c_void_p]),
    "ia_catalog_topk": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "ia_catalog_topk_seeded": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "ia_catalog_probe_bound": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "ia_catalog_last_stats": (c_int, [c_void_p, c_void_p]),
    "ia_catalog_last_plan": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ia_catalog_topk_dissimilarity": (c_int, [c_void_p, c_int, c_float, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "ia_catalog_file_write": (c_int, [c_char_p, c_int, c_void_p, c_int64, c_int64, c_void_p]),
    "ia_catalog_file_open": (c_int, [c_char_p, ctypes.POINTER(c_void_p)]),
    "ia_catalog_file_close": (None, [c_void_p]),
    "ia_catalog_file_info": (c_int, [c_void_p, ctypes.POINTER(c_int), ctypes.POINTER(c_int64), ctypes.POINTER(c_int64),
                                     ctypes.POINTER(c_int)]),
    "ia_catalog_file_data": (c_void_p, [c_void_p]),
    "ia_catalog_file_id": (c_void_p, [c_void_p, c_int64, ctypes.POINTER(c_int64)]),
    "ia_catalog_file_upload": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "ia_embedding_jsonl_to_catalog": (c_int, [c_char_p, c_char_p, c_int, c_int, c_int, ctypes.POINTER(c_int64),
                                              ctypes.POINTER(c_int64)]),
    "ia_topk_merge": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "ia_unpack_keys": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "ia_host_alloc": (c_int, [ctypes.POINTER(c_void_p), c_size_t, c_int]),
    "ia_host_free": (c_int, [c_void_p]),
    "ia_pair_score_host": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_double,
                                   c_void_p, c_int]),
    "ia_pair_score_loss_host": (c_int, [c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                        c_int64, c_void_p, c_void_p, c_void_p, c_int]),
}


class IAError(RuntimeError):
    pass


def lib():
    """Load the CUDA library once.  Raises (loudly) if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise IAError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C item_alignment_b200/csrc`).  item_alignment_b200 has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    """0 -> ok; IA_ERR_INVALID -> ValueError (the reference raises ValueError for a bad measure,
    src/models/base.py:64,86); everything else -> IAError."""
    if rc == 0:
        return
    msg = lib().ia_last_error().decode()
    if rc == -1:
        raise ValueError(msg)
    if rc == -2:
        raise NotImplementedError(msg)
    raise IAError(f"ia_b200 error {rc}: {msg}")


def launch_count():
    return int(lib().ia_launch_count())
