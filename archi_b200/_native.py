"""ctypes binding of libarchi_b200.so (include/archi_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, every entry point
raises.  Build it with ``python -m archi_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libarchi_b200.so")

# enums (include/archi_b200.h)
COSINE, L2, IP = 0, 1, 2
F32, BF16, I32, I64 = 0, 1, 2, 3
HOST, DEVICE = 0, 1
PATH_AUTO, PATH_STREAM, PATH_TENSOR = 0, 1, 2
METRICS = {"cosine": COSINE, "l2": L2, "inner_product": IP}
EUNSUPPORTED = -5

EXPORTS = [
    "archi_last_error", "archi_abi_version", "archi_kernel_launches",
    "archi_store_create", "archi_store_destroy", "archi_store_count", "archi_store_rows",
    "archi_store_info", "archi_store_reserve", "archi_store_reset", "archi_store_append",
    "archi_store_delete_rows", "archi_store_read_rows", "archi_store_save", "archi_store_load",
    "archi_pool_normalize", "archi_pool_normalize_append", "archi_search", "archi_hybrid_search",
    "archi_hybrid_search_terms", "archi_bm25_accumulate", "archi_merge_topk", "archi_merge_topk_strided", "archi_store_last_stats", "archi_store_set_timing",
    "archi_exchange_create", "archi_exchange_local_handle", "archi_exchange_connect", "archi_exchange_merge_topk",
    "archi_exchange_status", "archi_exchange_destroy",
]


class SearchStats(ctypes.Structure):
    _fields_ = [("path", ctypes.c_int), ("passes", ctypes.c_int), ("grid", ctypes.c_int),
                ("unverified_queries", ctypes.c_int), ("last_kernel_ms", ctypes.c_double),
                ("coarse_dtype", ctypes.c_int), ("coarse_launches", ctypes.c_int), ("unproven_queries", ctypes.c_int)]


class Bm25Terms(ctypes.Structure):
    """archi_bm25_terms_t (include/archi_b200.h)."""
    _fields_ = [("n_terms", ctypes.c_int), ("term_query", ctypes.c_void_p), ("post_start", ctypes.c_void_p),
                ("post_end", ctypes.c_void_p), ("idf", ctypes.c_void_p), ("doc_ids_dev", ctypes.c_void_p),
                ("tfs_dev", ctypes.c_void_p), ("doc_len_dev", ctypes.c_void_p), ("avgdl", ctypes.c_float),
                ("k1", ctypes.c_float), ("b", ctypes.c_float), ("sign", ctypes.c_float)]


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libarchi_b200: {message} (code {code})")
        self.code = code


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """Load the shared library once; raise loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library has not been built and archi_b200 has no "
            "CPU fallback.  Run `python -m archi_b200.build`.")
    L = ctypes.CDLL(LIB_PATH)
    c_i, c_i64, c_f, c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
    pp = ctypes.POINTER(ctypes.c_void_p)
    pi64 = ctypes.POINTER(ctypes.c_int64)
    pint = ctypes.POINTER(ctypes.c_int)
    L.archi_last_error.restype = ctypes.c_char_p
    L.archi_last_error.argtypes = []
    L.archi_abi_version.restype = c_i
    L.archi_kernel_launches.restype = c_i64
    L.archi_store_create.argtypes = [c_i, c_i, c_i, c_i, c_i64, pp]
    L.archi_store_destroy.argtypes = [c_p]
    L.archi_store_count.argtypes = [c_p, pi64]
    L.archi_store_rows.argtypes = [c_p, pi64]
    L.archi_store_info.argtypes = [c_p, pint, pint, pint, pint, pi64]
    L.archi_store_reserve.argtypes = [c_p, c_i64]
    L.archi_store_reset.argtypes = [c_p]
    L.archi_store_append.argtypes = [c_p, c_p, c_i, c_i, c_i64, c_p, pi64]
    L.archi_store_delete_rows.argtypes = [c_p, c_p, c_i64]
    L.archi_store_read_rows.argtypes = [c_p, c_i64, c_i64, c_p]
    L.archi_store_save.argtypes = [c_p, ctypes.c_char_p]
    L.archi_store_load.argtypes = [ctypes.c_char_p, c_i, pp]
    L.archi_pool_normalize.argtypes = [c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]
    L.archi_pool_normalize_append.argtypes = [c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_p, c_p, pi64]
    L.archi_search.argtypes = [c_p, c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_p, c_i, c_i64, c_p]
    L.archi_hybrid_search.argtypes = [c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_i, c_p, c_p, c_i,
                                      c_i64, c_p]
    L.archi_hybrid_search_terms.argtypes = [c_p, c_p, c_i, c_i, c_i, c_f, c_f, ctypes.POINTER(Bm25Terms), c_p, c_i, c_p, c_p,
                                            c_i, c_i64, c_p, pint]
    L.archi_bm25_accumulate.argtypes = [c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_f, c_f, c_f, c_f, c_p, c_p]
    L.archi_merge_topk.argtypes = [c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]
    L.archi_merge_topk_strided.argtypes = [c_i, c_p, c_p, c_i64, c_i64, c_i, c_i, c_i, c_i, c_p, c_p, c_p]
    L.archi_exchange_create.argtypes = [c_i, c_i, c_i, c_i64, pp]
    L.archi_exchange_local_handle.argtypes = [c_p, c_p]
    L.archi_exchange_connect.argtypes = [c_p, c_p]
    L.archi_exchange_merge_topk.argtypes = [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p]
    L.archi_exchange_status.argtypes = [c_p, pint]
    L.archi_exchange_destroy.argtypes = [c_p]
    L.archi_store_last_stats.argtypes = [c_p, ctypes.POINTER(SearchStats)]
    L.archi_store_set_timing.argtypes = [c_p, c_i]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("archi_last_error", "archi_abi_version", "archi_kernel_launches"):
            fn.restype = c_i
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise NativeError(rc, lib().archi_last_error().decode("utf-8", "replace"))


def kernel_launches() -> int:
    return int(lib().archi_kernel_launches())
