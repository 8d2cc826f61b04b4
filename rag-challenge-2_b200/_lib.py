"""ctypes binding of libb2r.so — the C ABI declared in include/b2r.h.

There is no CPU fallback: if the shared library is missing, or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2r.so")

B2R_OK = 0
B2R_ERR_INVALID = -1
B2R_ERR_CUDA = -2
B2R_ERR_NOMEM = -3
B2R_ERR_UNSUPPORTED = -4
B2R_MAX_K = 1024


class B2RError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libb2r error {code}: {message}")
        self.code = code


_c_f32p = C.POINTER(C.c_float)
_c_f64p = C.POINTER(C.c_double)
_c_i64p = C.POINTER(C.c_int64)
_c_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/b2r.h declares must be listed here
SIGNATURES = {
    "b2r_last_error": (C.c_char_p, []),
    "b2r_version": (C.c_int, []),
    "b2r_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b2r_tunable_set": (C.c_int, [C.c_char_p, C.c_int]),
    "b2r_tunable_get": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "b2r_flat_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "b2r_flat_add": (C.c_int, [_vp, _vp, C.c_int64, C.c_int]),
    "b2r_flat_adopt": (C.c_int, [_vp, _vp, C.c_int64]),
    "b2r_flat_ntotal": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "b2r_flat_dim": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "b2r_flat_rows_dev": (C.c_int, [_vp, C.POINTER(_vp)]),
    "b2r_flat_search_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "b2r_flat_search_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, _vp]),
    "b2r_flat_combine_stats": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b2r_flat_search_dev_ex": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, _vp, C.c_uint]),
    "b2r_flat_reconstruct": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "b2r_flat_set_candidate_bf16": (C.c_int, [_vp, C.c_int]),
    "b2r_flat_destroy": (C.c_int, [_vp]),
    "b2r_last_launch_count": (C.c_int, []),
    "b2r_topk_merge_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "b2r_topk_merge_f64_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "b2r_hybrid_fuse_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "b2r_hybrid_exchange_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "b2r_hybrid_exchange_fuse_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(_vp), C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_uint32, _vp, _vp, _vp, _vp]),
    "b2r_bm25_build_count_dev": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                          _vp]),
    "b2r_bm25_build_place_dev": (C.c_int, [_vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2r_cosine_heatmap_dev": (C.c_int, [_vp, C.c_int64, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "b2r_hybrid_step_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int64, C.POINTER(_vp), C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_uint,
                                     C.POINTER(C.c_int)]),
    "b2r_hybrid_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "b2r_hybrid_ctx_destroy": (C.c_int, [_vp]),
    "b2r_hybrid_ctx_bytes": (C.c_int, [_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "b2r_hybrid_search_host": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int64,
                                        C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, _vp, _vp]),
    "b2r_hybrid_submit_host": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int64,
                                        C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_int)]),
    "b2r_hybrid_collect_host": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "b2r_topk_merge_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int]),
    "b2r_bm25_create": (C.c_int, [C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, C.c_double, _vp, C.c_double,
                                  C.c_double, C.c_int, C.POINTER(_vp)]),
    "b2r_bm25_create_ex": (C.c_int, [C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, C.c_double, _vp, C.c_double,
                                     C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "b2r_bm25_ndocs": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "b2r_bm25_postings": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "b2r_bm25_scores_host": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "b2r_bm25_topk_host": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "b2r_bm25_topk_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, _vp]),
    "b2r_bm25_topk_dev_ex": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, _vp, C.c_uint]),
    "b2r_bm25_destroy": (C.c_int, [_vp]),
    "b2r_cosine_stats_dev": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int64, C.c_int64, _vp, _vp, C.c_int, C.c_int,
                                       _vp, _vp, C.c_int, _vp]),
}

_lib = None


def load():
    """Load libb2r.so (once) and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run ./build.sh (or __graft_entry__.build()). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != B2R_OK:
        msg = load().b2r_last_error()
        raise B2RError(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc


def tunable_set(key, value):
    check(load().b2r_tunable_set(key.encode(), int(value)))


def tunable_get(key):
    v = C.c_int(0)
    check(load().b2r_tunable_get(key.encode(), C.byref(v)))
    return v.value


def device_count():
    v = C.c_int(0)
    check(load().b2r_device_count(C.byref(v)))
    return v.value


def ptr(a):
    """void* of a numpy array (host) — the array must stay alive during the call."""
    return a.ctypes.data_as(_vp)
