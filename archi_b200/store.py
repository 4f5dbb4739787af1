"""NativeStore -- Python handle on one GPU-resident row shard (archi_store_t in include/archi_b200.h).

This is the thin layer between the reference-facing classes (vectorstore.py, sharded.py,
embeddings.py) and the C ABI.  numpy arrays are passed as HOST buffers, torch CUDA tensors as
DEVICE buffers; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _native as N


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _current_stream_ptr(device: int) -> int:
    import torch
    return int(torch.cuda.current_stream(device).cuda_stream)


def _torch_dtype_code(t) -> int:
    import torch
    if t.dtype == torch.float32:
        return N.F32
    if t.dtype == torch.bfloat16:
        return N.BF16
    if t.dtype == torch.int32:
        return N.I32
    if t.dtype == torch.int64:
        return N.I64
    raise TypeError(f"unsupported tensor dtype {t.dtype}")


class NativeStore:
    """One row shard in the HBM of ``device``; dim and metric are fixed at creation."""

    def __init__(self, dim: int, metric: str = "cosine", storage_dtype: str = "f32", device: int = 0,
                 capacity_rows: int = 0, _handle: Optional[int] = None):
        if metric not in N.METRICS:
            raise ValueError(f"distance_metric must be one of {list(N.METRICS.keys())}")
        if storage_dtype not in ("f32", "bf16"):
            raise ValueError("storage_dtype must be 'f32' or 'bf16'")
        self.dim, self.metric, self.storage_dtype, self.device = int(dim), metric, storage_dtype, int(device)
        self._L = N.lib()
        self.last_hybrid_path = "none"
        if _handle is None:
            h = ctypes.c_void_p()
            N.check(self._L.archi_store_create(self.device, self.dim, N.METRICS[metric],
                                               N.BF16 if storage_dtype == "bf16" else N.F32,
                                               int(capacity_rows), ctypes.byref(h)))
            self._h = h
        else:
            self._h = ctypes.c_void_p(_handle)

    # ---- lifetime ------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._L.archi_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def larger_is_better(self) -> bool:
        """Output-score direction: cosine similarity is larger-better; l2 and inner_product return
        distances (postgres_vectorstore.py:361)."""
        return self.metric == "cosine"

    # ---- bookkeeping -----------------------------------------------------------------------------
    def count(self) -> int:
        out = ctypes.c_int64()
        N.check(self._L.archi_store_count(self._h, ctypes.byref(out)))
        return int(out.value)

    def rows(self) -> int:
        out = ctypes.c_int64()
        N.check(self._L.archi_store_rows(self._h, ctypes.byref(out)))
        return int(out.value)

    def capacity(self) -> int:
        cap = ctypes.c_int64()
        N.check(self._L.archi_store_info(self._h, None, None, None, None, ctypes.byref(cap)))
        return int(cap.value)

    def reserve(self, capacity_rows: int) -> None:
        N.check(self._L.archi_store_reserve(self._h, int(capacity_rows)))

    def reset(self) -> None:
        N.check(self._L.archi_store_reset(self._h))

    # ---- rows ------------------------------------------------------------------------------------
    def append(self, rows) -> int:
        """Append [n, dim] rows (numpy fp32 on the host, or a torch CUDA tensor fp32/bf16).
        Returns the id of the first appended row."""
        first = ctypes.c_int64()
        if _is_torch_tensor(rows):
            t = rows.contiguous()
            if t.dim() != 2 or t.shape[1] != self.dim:
                raise ValueError(f"expected [n, {self.dim}] rows, got {tuple(t.shape)}")
            if not t.is_cuda or t.device.index != self.device:
                raise ValueError("device rows must live on the store's GPU")
            N.check(self._L.archi_store_append(self._h, ctypes.c_void_p(t.data_ptr()), _torch_dtype_code(t),
                                               N.DEVICE, t.shape[0], ctypes.c_void_p(_current_stream_ptr(self.device)),
                                               ctypes.byref(first)))
        else:
            a = np.ascontiguousarray(rows, dtype=np.float32)
            if a.ndim != 2 or a.shape[1] != self.dim:
                raise ValueError(f"expected [n, {self.dim}] rows, got {a.shape}")
            N.check(self._L.archi_store_append(self._h, a.ctypes.data_as(ctypes.c_void_p), N.F32, N.HOST,
                                               a.shape[0], None, ctypes.byref(first)))
        return int(first.value)

    def delete_rows(self, rows) -> None:
        a = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        N.check(self._L.archi_store_delete_rows(self._h, a.ctypes.data_as(ctypes.c_void_p), a.size))

    def read_rows(self, first_row: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=np.float32)
        N.check(self._L.archi_store_read_rows(self._h, int(first_row), int(n), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def save(self, path: str) -> None:
        N.check(self._L.archi_store_save(self._h, path.encode()))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "NativeStore":
        L = N.lib()
        h = ctypes.c_void_p()
        N.check(L.archi_store_load(path.encode(), int(device), ctypes.byref(h)))
        dim, metric, dtype = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        N.check(L.archi_store_info(h, ctypes.byref(dim), ctypes.byref(metric), ctypes.byref(dtype), None, None))
        name = {v: k for k, v in N.METRICS.items()}[metric.value]
        return cls(dim.value, name, "bf16" if dtype.value == N.BF16 else "f32", device, _handle=h.value)

    # ---- pool + normalise fused with the append -----------------------------------------------------
    def pool_normalize_append(self, hidden, mask, out_f32=None) -> int:
        """hidden [B, L, dim] (torch CUDA fp32/bf16), mask [B, L] (int32/int64): one kernel writes
        the B normalised rows into the store's tail.  Returns the first row id."""
        h, m = hidden.contiguous(), mask.contiguous()
        B, Lq, H = h.shape
        if H != self.dim:
            raise ValueError(f"hidden size {H} != store dim {self.dim}")
        first = ctypes.c_int64()
        N.check(self._L.archi_pool_normalize_append(
            self._h, ctypes.c_void_p(h.data_ptr()), _torch_dtype_code(h), ctypes.c_void_p(m.data_ptr()),
            _torch_dtype_code(m), B, Lq, ctypes.c_void_p(out_f32.data_ptr()) if out_f32 is not None else None,
            ctypes.c_void_p(_current_stream_ptr(self.device)), ctypes.byref(first)))
        return int(first.value)

    # ---- search --------------------------------------------------------------------------------------
    def search(self, queries, k: int, filter_mask=None, include_deleted: bool = False, id_offset: int = 0,
               path: int = N.PATH_AUTO, bm25=None, semantic_weight: float = 1.0, bm25_weight: float = 0.0,
               hybrid: bool = False, out=None):
        """Exact top-k.  numpy queries -> numpy (scores [nq,k] fp32, ids [nq,k] int64) through host
        buffers; torch CUDA queries -> torch CUDA outputs, nothing synchronised.
        ``filter_mask``: torch CUDA int32/uint32 bitmask words (bit i&31 of word i>>5 = row passes).
        ``hybrid``: scores are the combined score of hybrid_search; ``bm25`` is a torch CUDA fp32
        [nq, rows] tensor (or None).  ``out``: optional preallocated (scores, ids) -- host arrays
        (e.g. views of pinned memory) on the numpy path, contiguous CUDA tensors on the torch path."""
        k = int(k)
        fm = ctypes.c_void_p(filter_mask.data_ptr()) if filter_mask is not None else None
        bm = ctypes.c_void_p(bm25.data_ptr()) if bm25 is not None else None
        if _is_torch_tensor(queries):
            import torch
            q = queries.contiguous().to(torch.float32)
            if q.dim() == 1:
                q = q[None, :]
            nq = q.shape[0]
            if out is not None:
                scores, ids = out
                if tuple(scores.shape) != (nq, k) or tuple(ids.shape) != (nq, k) or scores.dtype != torch.float32 \
                        or ids.dtype != torch.int64 or not (scores.is_contiguous() and ids.is_contiguous()) \
                        or scores.device != q.device or ids.device != q.device:
                    raise ValueError("out must be contiguous CUDA tensors (float32 [nq,k], int64 [nq,k]) on the query device")
            else:
                scores = torch.empty((nq, k), dtype=torch.float32, device=q.device)
                ids = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            qp, sp, ip_, loc = q.data_ptr(), scores.data_ptr(), ids.data_ptr(), N.DEVICE
            stream = ctypes.c_void_p(_current_stream_ptr(self.device))
        else:
            q = np.ascontiguousarray(np.atleast_2d(np.asarray(queries, dtype=np.float32)))
            nq = q.shape[0]
            if out is not None:
                scores, ids = out
                if scores.shape != (nq, k) or ids.shape != (nq, k) or scores.dtype != np.float32 \
                        or ids.dtype != np.int64 or not (scores.flags.c_contiguous and ids.flags.c_contiguous):
                    raise ValueError("out must be C-contiguous (float32 [nq,k], int64 [nq,k])")
            else:
                scores = np.empty((nq, k), dtype=np.float32)
                ids = np.empty((nq, k), dtype=np.int64)
            qp, sp, ip_, loc = q.ctypes.data, scores.ctypes.data, ids.ctypes.data, N.HOST
            stream = ctypes.c_void_p(_current_stream_ptr(self.device)) if (fm or bm) else None
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != store dimension {self.dim}")
        if hybrid:
            N.check(self._L.archi_hybrid_search(self._h, ctypes.c_void_p(qp), loc, nq, k, float(semantic_weight),
                                                float(bm25_weight), bm, fm, int(include_deleted),
                                                ctypes.c_void_p(sp), ctypes.c_void_p(ip_), loc, int(id_offset), stream))
        else:
            N.check(self._L.archi_search(self._h, ctypes.c_void_p(qp), loc, nq, k, fm, int(include_deleted), int(path),
                                         ctypes.c_void_p(sp), ctypes.c_void_p(ip_), loc, int(id_offset), stream))
        return scores, ids

    def hybrid_search_terms(self, lexical, text_queries, queries, k: int, semantic_weight: float, bm25_weight: float,
                            filter_mask=None, include_deleted: bool = False, id_offset: int = 0):
        """hybrid_search over posting lists (archi_hybrid_search_terms): ``lexical`` is the collection's
        LexicalIndex, ``text_queries`` one lexical query per embedding in ``queries`` (a string, or an array of
        term keys).  numpy embeddings -> numpy outputs, torch CUDA embeddings -> torch CUDA outputs (nothing
        synchronised).  Scores are the combined scores, best first.  ``last_hybrid_path`` says which path ran."""
        k = int(k)
        ranges = [lexical.posting_ranges(tq) for tq in text_queries]
        counts = [r[0].size for r in ranges]
        term_query = np.repeat(np.arange(len(ranges), dtype=np.int32), counts)
        starts = np.concatenate([r[0] for r in ranges]) if ranges else np.empty(0, np.int64)
        ends = np.concatenate([r[1] for r in ranges]) if ranges else np.empty(0, np.int64)
        idf = np.concatenate([r[2] for r in ranges]) if ranges else np.empty(0, np.float32)
        doc_ids, tfs, doc_len = lexical.device_arrays()
        terms = N.Bm25Terms(int(term_query.size), term_query.ctypes.data, starts.ctypes.data, ends.ctypes.data, idf.ctypes.data,
                            doc_ids.data_ptr(), tfs.data_ptr(), doc_len.data_ptr(), float(lexical.avgdl()),
                            float(lexical.k1), float(lexical.b), float(lexical.sign))
        fm = ctypes.c_void_p(filter_mask.data_ptr()) if filter_mask is not None else None
        stream = ctypes.c_void_p(_current_stream_ptr(self.device))
        if _is_torch_tensor(queries):
            import torch
            q = queries.contiguous().to(torch.float32)
            if q.dim() == 1:
                q = q[None, :]
            nq = q.shape[0]
            scores = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            ids = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            qp, sp, ip_, loc = q.data_ptr(), scores.data_ptr(), ids.data_ptr(), N.DEVICE
        else:
            q = np.ascontiguousarray(np.atleast_2d(np.asarray(queries, dtype=np.float32)))
            nq = q.shape[0]
            scores = np.empty((nq, k), dtype=np.float32)
            ids = np.empty((nq, k), dtype=np.int64)
            qp, sp, ip_, loc = q.ctypes.data, scores.ctypes.data, ids.ctypes.data, N.HOST
        if nq != len(ranges):
            raise ValueError("one lexical query per embedding is required")
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != store dimension {self.dim}")
        path = ctypes.c_int(0)
        N.check(self._L.archi_hybrid_search_terms(self._h, ctypes.c_void_p(qp), loc, nq, k, float(semantic_weight),
                                                  float(bm25_weight), ctypes.byref(terms), fm, int(include_deleted),
                                                  ctypes.c_void_p(sp), ctypes.c_void_p(ip_), loc, int(id_offset), stream,
                                                  ctypes.byref(path)))
        self.last_hybrid_path = {1: "posting-lists", 2: "dense-vector"}.get(path.value, "none")
        return scores, ids

    def last_stats(self) -> N.SearchStats:
        st = N.SearchStats()
        N.check(self._L.archi_store_last_stats(self._h, ctypes.byref(st)))
        return st

    def set_timing(self, enabled: bool) -> None:
        N.check(self._L.archi_store_set_timing(self._h, int(bool(enabled))))


def pool_normalize(hidden, mask, want_bf16: bool = False, want_f32: bool = True):
    """Fused masked mean-pool + L2 normalise (+ bf16 cast) of ``hidden`` [B, L, H] on its GPU."""
    import torch
    h, m = hidden.contiguous(), mask.contiguous()
    B, Lq, H = h.shape
    dev = h.device
    out_f32 = torch.empty((B, H), dtype=torch.float32, device=dev) if want_f32 else None
    out_bf16 = torch.empty((B, H), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    with torch.cuda.device(dev):
        N.check(N.lib().archi_pool_normalize(
            ctypes.c_void_p(h.data_ptr()), _torch_dtype_code(h), ctypes.c_void_p(m.data_ptr()), _torch_dtype_code(m),
            B, Lq, H, ctypes.c_void_p(out_bf16.data_ptr()) if want_bf16 else None,
            ctypes.c_void_p(out_f32.data_ptr()) if want_f32 else None,
            ctypes.c_void_p(_current_stream_ptr(dev.index))))
    return out_f32, out_bf16


def merge_topk(scores, ids, larger_is_better: bool):
    """scores/ids: torch CUDA [n_lists, nq, k] (each list best first) -> merged [nq, k]."""
    import torch

    def list_strided(t):
        # dense [nq, k] blocks, any distance between lists (views into one all-gather buffer)
        return t.dim() == 3 and (t.shape[1] * t.shape[2] == 0 or
                                 (t.stride(2) == 1 and t.stride(1) == t.shape[2] and t.stride(0) >= t.shape[1] * t.shape[2]))

    s = scores if list_strided(scores) else scores.contiguous()
    i = ids if list_strided(ids) else ids.contiguous()
    n_lists, nq, k = s.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=s.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=s.device)
    N.check(N.lib().archi_merge_topk_strided(
        s.device.index, ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(i.data_ptr()),
        int(s.stride(0)) if n_lists > 1 else nq * k, int(i.stride(0)) if n_lists > 1 else nq * k,
        n_lists, nq, k, int(larger_is_better), ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
        ctypes.c_void_p(_current_stream_ptr(s.device.index))))
    return out_s, out_i
