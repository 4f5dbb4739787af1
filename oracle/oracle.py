"""
oracle.py -- numpy restatement of archi's retrieval hot path + ctypes loader for oracle.c.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under archi_b200/ may import this module.

PARITY UNPINNED (see oracle.c header and DESIGN.md): pgvector, pg_textsearch and
sentence-transformers hold the arithmetic and are absent from /root/reference and from this image.
Their published algorithms are restated; the reference's own call sites and score conventions
(postgres_vectorstore.py:74-78, 317-332, 361, 441-456, 466-469) are what the restatement follows.

Two precision levels:
  * ``exact_topk`` / ``exact_scores``: fp64 accumulation over the values *as stored* (bf16 stores
    are up-cast exactly).  This is the "fp32 exact search" truth the CUDA path is compared with
    (ids identical up to tie order; scores within 1e-5 relative for fp32 storage, 2e-3 for bf16).
  * oracle.c (``clib()``): float accumulators in index order -- the reference CPU path's own
    precision, used for the cpu_baseline timing and as a second checker.
"""
from __future__ import annotations

import ctypes
import math
import os
import re
import subprocess
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = {"cosine": 0, "l2": 1, "inner_product": 2}
# postgres_vectorstore.py:74-78
DISTANCE_OPS = {"cosine": "<=>", "l2": "<->", "inner_product": "<#>"}


# --------------------------------------------------------------------------------------------
# bf16 helpers (numpy has no bfloat16): storage is uint16 holding the top half of an fp32.
# --------------------------------------------------------------------------------------------
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 bit pattern (what torch .to(bfloat16) / __float2bfloat16_rn do)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    return (rounded >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


# --------------------------------------------------------------------------------------------
# Distances and the reference's score conventions
# --------------------------------------------------------------------------------------------
def distances_f64(metric: str, corpus: np.ndarray, queries: np.ndarray) -> np.ndarray:
    """[Q, N] fp64 distances with pgvector's operator definitions [external]:
    cosine ``<=>`` = 1 - a.b/sqrt(|a|^2 |b|^2) (similarity clamped to [-1,1]),
    l2 ``<->`` = sqrt(sum (a-b)^2), inner_product ``<#>`` = -a.b.  Only for small N*Q."""
    c = np.asarray(corpus, dtype=np.float64)
    q = np.atleast_2d(np.asarray(queries, dtype=np.float64))
    dot = q @ c.T
    if metric == "inner_product":
        return -dot
    if metric == "l2":
        d2 = (q * q).sum(1)[:, None] + (c * c).sum(1)[None, :] - 2.0 * dot
        # the expansion cancels badly for near-identical vectors; recompute those directly
        small = d2 < 1e-6 * ((q * q).sum(1)[:, None] + (c * c).sum(1)[None, :])
        if small.any():
            qi, ci = np.nonzero(small)
            d2[qi, ci] = ((q[qi] - c[ci]) ** 2).sum(1)
        return np.sqrt(np.maximum(d2, 0.0))
    if metric == "cosine":
        with np.errstate(divide="ignore", invalid="ignore"):
            sim = dot / np.sqrt((q * q).sum(1)[:, None] * (c * c).sum(1)[None, :])
        sim = np.clip(sim, -1.0, 1.0)
        return 1.0 - sim
    raise ValueError(f"distance_metric must be one of {list(METRICS)}")


def score_from_distance(metric: str, distance):
    """postgres_vectorstore.py:361 -- ``1 - distance`` for cosine, the raw distance otherwise
    (so l2 returns a distance and inner_product returns the NEGATIVE inner product)."""
    return 1.0 - distance if metric == "cosine" else distance


def topk_from_keys(keys: np.ndarray, k: int, ascending: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Stable top-k along the last axis: best first, ties broken by lower id.  keys [Q, N]."""
    keys = np.atleast_2d(keys)
    n = keys.shape[1]
    kk = min(k, n)
    order_keys = keys if ascending else -keys
    if kk > 0 and n > 16 * kk and not np.isnan(order_keys).any():
        # same result as the full stable sort, without sorting N keys per query: everything up to the
        # kk-th smallest key (ties included) is a candidate, candidates are visited in id order
        kth = np.partition(order_keys, kk - 1, axis=1)[:, kk - 1]
        idx = np.empty((keys.shape[0], kk), dtype=np.int64)
        for r in range(keys.shape[0]):
            cand = np.flatnonzero(order_keys[r] <= kth[r])
            idx[r] = cand[np.argsort(order_keys[r, cand], kind="stable")[:kk]]
    else:
        idx = np.argsort(order_keys, axis=1, kind="stable")[:, :kk]
    vals = np.take_along_axis(keys, idx, axis=1)
    return vals, idx.astype(np.int64)


def exact_topk(metric: str, corpus: np.ndarray, queries: np.ndarray, k: int,
               mask: Optional[np.ndarray] = None, block: int = 262144):
    """fp64 exact search: (distances [Q,k'] ascending, ids [Q,k'] int64), k' = min(k, rows passing).
    ``mask`` is a boolean [N] array (True = row passes the WHERE clause).  Blocked over N so a
    1M x 384 corpus with a few queries stays in memory."""
    corpus = np.asarray(corpus)
    q = np.atleast_2d(np.asarray(queries, dtype=np.float64))
    n = corpus.shape[0]
    best_d = np.empty((q.shape[0], 0))
    best_i = np.empty((q.shape[0], 0), dtype=np.int64)
    for s in range(0, n, block):
        e = min(n, s + block)
        d = distances_f64(metric, corpus[s:e], q)
        d = np.where(np.isnan(d), np.inf, d)
        ids = np.arange(s, e, dtype=np.int64)
        if mask is not None:
            keep = np.asarray(mask[s:e], dtype=bool)
            d = d[:, keep]
            ids = ids[keep]
        if d.shape[1] == 0:
            continue
        vals, idx = topk_from_keys(d, k)
        cand_d = np.concatenate([best_d, vals], axis=1)
        cand_i = np.concatenate([best_i, np.broadcast_to(ids, (q.shape[0], ids.size))[
            np.arange(q.shape[0])[:, None], idx]], axis=1)
        # ids are increasing across blocks, so a stable sort keeps "lower id wins"
        o = np.argsort(cand_d, axis=1, kind="stable")[:, :k]
        best_d = np.take_along_axis(cand_d, o, axis=1)
        best_i = np.take_along_axis(cand_i, o, axis=1)
    return best_d, best_i


def exact_hybrid_topk(metric: str, corpus: np.ndarray, query: np.ndarray, bm25: Optional[np.ndarray],
                      semantic_weight: float, bm25_weight: float, k: int,
                      mask: Optional[np.ndarray] = None):
    """postgres_vectorstore.py:441-456 in fp64.  ``bm25`` is a dense [N] array with NaN = SQL NULL.
    Returns (combined [k'] descending, ids [k'])."""
    d = distances_f64(metric, corpus, query)[0]
    sem = 1.0 - d
    b = np.zeros_like(sem) if bm25 is None else np.where(np.isnan(bm25), 0.0, bm25)
    combined = sem * semantic_weight + b * bm25_weight
    ids = np.arange(corpus.shape[0], dtype=np.int64)
    if mask is not None:
        keep = np.asarray(mask, dtype=bool)
        combined, ids = combined[keep], ids[keep]
    if combined.size == 0:
        return combined, ids
    vals, idx = topk_from_keys(combined[None, :], k, ascending=False)
    return vals[0], ids[idx[0]]


def same_topk_up_to_ties(ids_a: Sequence[int], ids_b: Sequence[int], keys_b: Sequence[float],
                         rel_tol: float = 0.0, abs_tol: float = 0.0) -> bool:
    """True when the id sets agree except among rows whose truth key ties the k-th key
    (``keys_b`` are the truth keys of ``ids_b``, best first).  With tolerances > 0 a row whose
    truth key is within tol of the k-th key counts as tied -- used for bf16/approximate paths."""
    a, b = list(map(int, ids_a)), list(map(int, ids_b))
    if len(a) != len(b):
        return False
    if set(a) == set(b):
        return True
    kth = float(keys_b[-1])
    tol = abs_tol + rel_tol * abs(kth)
    strict = {i for i, s in zip(b, keys_b) if abs(float(s) - kth) > tol}
    return strict.issubset(set(a))


def pair_distances_f64(metric: str, rows: np.ndarray, query: np.ndarray) -> np.ndarray:
    """fp64 distances of ONE query to a handful of gathered rows [m, D] (same definitions as distances_f64)."""
    if rows.shape[0] == 0:
        return np.empty(0)
    d = distances_f64(metric, rows, query)[0]
    return np.where(np.isnan(d), np.inf, d)


def verify_topk(metric: str, stored: np.ndarray, queries: np.ndarray, k: int, ids: np.ndarray, scores: np.ndarray,
                rel: float, d_true: np.ndarray, i_true: np.ndarray, mask: Optional[np.ndarray] = None,
                id_tol_rel: float = 2e-6, id_tol_abs: float = 1e-7) -> List[str]:
    """Tie-aware and exact comparison of a returned top-k with the fp64 truth (d_true, i_true of exact_topk or
    oracle.c on the same stored values).  Unlike ``same_topk_up_to_ties`` every returned id is re-scored in
    fp64 from ``stored``, so an id outside the truth list is accepted only if its own distance ties the k-th
    truth distance.  Returns a list of failure descriptions (empty = pass).  Per query:
      * min(k, rows passing) distinct ids, all passing ``mask``; the tail is id -1 / score NaN;
      * every truth id strictly better than the k-th truth distance (beyond the tie tolerance) is returned;
      * every returned id's fp64 distance is <= the k-th truth distance + tolerance;
      * returned scores, rank by rank, within ``rel`` of the truth scores and best first."""
    fails: List[str] = []
    queries = np.atleast_2d(queries)
    s_true = score_from_distance(metric, d_true)
    kk = d_true.shape[1]
    for q in range(queries.shape[0]):
        got = [int(x) for x in ids[q, :kk]]
        if len(set(got)) != kk or min(got, default=0) < 0:
            fails.append(f"q{q}: ids not distinct/valid {got}")
            continue
        if not ((ids[q, kk:] == -1).all() and np.isnan(scores[q, kk:]).all()):
            fails.append(f"q{q}: tail not empty")
        if mask is not None and not np.asarray(mask)[got].all():
            fails.append(f"q{q}: returned a masked row")
        if kk == 0:
            continue
        kth = float(d_true[q, -1])
        tol = id_tol_abs + id_tol_rel * abs(kth)
        strict = {int(i) for i, d in zip(i_true[q], d_true[q]) if d < kth - tol}
        if not strict.issubset(got):
            fails.append(f"q{q}: missing strictly-better ids {sorted(strict - set(got))}")
        extra = [i for i in got if i not in set(int(x) for x in i_true[q])]
        if extra:
            d_extra = pair_distances_f64(metric, np.asarray(stored[extra], dtype=np.float64), queries[q])
            if (d_extra > kth + tol).any():
                fails.append(f"q{q}: ids {extra} are not ties of the k-th distance {kth}: {d_extra.tolist()}")
        if not np.allclose(scores[q, :kk], s_true[q], rtol=rel, atol=rel * 1e-1):
            fails.append(f"q{q}: scores {scores[q, :kk].tolist()} vs {s_true[q].tolist()}")
        key = scores[q, :kk] if metric == "cosine" else -scores[q, :kk]
        if not (np.diff(key) <= 1e-7).all():
            fails.append(f"q{q}: not best first")
    return fails


# --------------------------------------------------------------------------------------------
# Pool + normalise (sentence-transformers Pooling(mean) + Normalize [external])
# --------------------------------------------------------------------------------------------
def pool_normalize(hidden: np.ndarray, mask: np.ndarray, dtype=np.float64) -> np.ndarray:
    """sum_t h*m / max(sum_t m, 1e-9), then x / max(|x|_2, 1e-12).  hidden [B,L,H], mask [B,L]."""
    h = np.asarray(hidden, dtype=dtype)
    m = np.asarray(mask).astype(dtype)
    pooled = (h * m[:, :, None]).sum(1) / np.maximum(m.sum(1), 1e-9)[:, None]
    nrm = np.sqrt((pooled * pooled).sum(1))
    return pooled / np.maximum(nrm, 1e-12)[:, None]


# --------------------------------------------------------------------------------------------
# BM25 restatement (pg_textsearch 0.4.2 ``<@>`` [external, unverified]) -- parity unpinned.
# Contract taken from the reference's own tests and docs: bm25 >= 0, higher is better
# (test_postgres_vectorstore.py:259-261,304-306; docs/docs/configuration.md:155).  The literal
# upstream operator is documented to return the NEGATED score; ``sign=-1`` reproduces that.
# --------------------------------------------------------------------------------------------
_TOKEN_RE = re.compile(r"[a-z0-9]+")


def tokenize(text: str) -> List[str]:
    """Lower-case alphanumeric runs.  (Postgres' english config also stems and drops stop
    words; not restated -- recorded as an assumption in DESIGN.md.)"""
    return _TOKEN_RE.findall(text.lower())


def bm25_scores(docs_tokens: Sequence[Sequence[str]], query_tokens: Sequence[str],
                k1: float = 1.2, b: float = 0.75, sign: float = 1.0) -> np.ndarray:
    """Dense [N] fp64 BM25 with Lucene-style idf = ln(1 + (N - df + 0.5)/(df + 0.5)); rows that
    share no term with the query get NaN (SQL NULL -> COALESCE(.,0))."""
    n = len(docs_tokens)
    out = np.full(n, np.nan)
    if n == 0:
        return out
    dl = np.array([len(t) for t in docs_tokens], dtype=np.float64)
    avgdl = dl.mean() if dl.sum() > 0 else 1.0
    tfs: List[Dict[str, int]] = []
    df: Dict[str, int] = {}
    for toks in docs_tokens:
        c: Dict[str, int] = {}
        for t in toks:
            c[t] = c.get(t, 0) + 1
        tfs.append(c)
        for t in c:
            df[t] = df.get(t, 0) + 1
    # repeated query terms count once per occurrence, as a bag-of-words query does
    for i, c in enumerate(tfs):
        s, hit = 0.0, False
        for t in query_tokens:
            tf = c.get(t, 0)
            if tf == 0:
                continue
            hit = True
            idf = math.log(1.0 + (n - df[t] + 0.5) / (df[t] + 0.5))
            s += idf * tf * (k1 + 1.0) / (tf + k1 * (1.0 - b + b * dl[i] / avgdl))
        if hit:
            out[i] = sign * s
    return out


# --------------------------------------------------------------------------------------------
# Chunking for config 1 (manager.py:75-78,292: CharacterTextSplitter("\n\n", 1000, 0) [external])
# --------------------------------------------------------------------------------------------
def character_text_split(text: str, chunk_size: int = 1000, chunk_overlap: int = 0,
                         separator: str = "\n\n") -> List[str]:
    """Split on the separator, then greedily merge pieces up to chunk_size characters (pieces
    longer than chunk_size are kept whole -- the upstream splitter only warns)."""
    pieces = [p for p in text.split(separator) if p != ""]
    chunks: List[str] = []
    cur: List[str] = []
    total = 0
    sep_len = len(separator)
    for p in pieces:
        extra = len(p) + (sep_len if cur else 0)
        if total + extra > chunk_size and cur:
            doc = separator.join(cur).strip()
            if doc:
                chunks.append(doc)
            while total > chunk_overlap or (total + extra > chunk_size and total > 0):
                total -= len(cur[0]) + (sep_len if len(cur) > 1 else 0)
                cur.pop(0)
                if not cur:
                    total = 0
                    break
        cur.append(p)
        total += len(p) + (sep_len if len(cur) > 1 else 0)
    doc = separator.join(cur).strip()
    if doc:
        chunks.append(doc)
    return chunks


# --------------------------------------------------------------------------------------------
# oracle.c via ctypes
# --------------------------------------------------------------------------------------------
_LIBS = {}

# pgvector's own build flags [external]: its Makefile compiles the distance loops with
# -march=native -ftree-vectorize -fassociative-math -fno-signed-zeros -fno-trapping-math, i.e. the
# float accumulators are SIMD-reassociated.  The strict build (fast=False) keeps index order and is
# the checker; the fast build is only the TIMED cpu baseline.
FAST_FLAGS = ["-O3", "-march=native", "-ftree-vectorize", "-fassociative-math", "-fno-signed-zeros",
              "-fno-trapping-math"]


def _cpu_tag() -> str:
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        key = "".join(l for l in txt.splitlines() if l.startswith(("model name", "flags")))[:20000]
    except Exception:
        key = "unknown"
    return hashlib.sha1(key.encode()).hexdigest()[:10]


def build_clib(force: bool = False, fast: bool = False) -> str:
    # the -march=native build is tagged with the host CPU so a copy built elsewhere is never loaded
    so = os.path.join(_HERE, f"liboracle_fast_{_cpu_tag()}.so" if fast else "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        flags = FAST_FLAGS if fast else ["-O2"]
        subprocess.check_call(["gcc"] + flags + ["-fPIC", "-shared", "-pthread", "-o", so, src, "-lm"])
    return so


def clib(fast: bool = False) -> ctypes.CDLL:
    if fast not in _LIBS:
        lib = ctypes.CDLL(build_clib(fast=fast))
        c_i, c_i64, c_d, c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
        lib.orc_distance_f32.restype = c_d
        lib.orc_distance_f32.argtypes = [c_i, c_i, c_p, c_p]
        lib.orc_scan_topk.restype = c_i
        lib.orc_scan_topk.argtypes = [c_i, c_p, c_i, c_i64, c_i, c_p, c_i, c_p, c_p, c_p]
        lib.orc_scan_topk_batch.restype = c_i
        lib.orc_scan_topk_batch.argtypes = [c_i, c_p, c_i, c_i64, c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_i]
        lib.orc_hybrid_topk.restype = c_i
        lib.orc_hybrid_topk.argtypes = [c_i, c_p, c_i, c_i64, c_i, c_p, c_p, c_d, c_d, c_i, c_p, c_p, c_p]
        lib.orc_score_from_distance.restype = c_d
        lib.orc_score_from_distance.argtypes = [c_i, c_d]
        lib.orc_pool_normalize.restype = None
        lib.orc_pool_normalize.argtypes = [c_p, c_p, c_i, c_i, c_i, c_p]
        _LIBS[fast] = lib
    return _LIBS[fast]


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def pack_mask(mask_bool: Optional[np.ndarray]) -> Optional[np.ndarray]:
    """bool [N] -> uint32 words, bit i of word i>>5 = row i (the layout the C-ABI also uses)."""
    if mask_bool is None:
        return None
    m = np.asarray(mask_bool, dtype=bool)
    pad = (-m.size) % 32
    if pad:
        m = np.concatenate([m, np.zeros(pad, dtype=bool)])
    return np.packbits(m.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1).copy()


def c_scan_topk(metric: str, corpus: np.ndarray, queries: np.ndarray, k: int,
                mask: Optional[np.ndarray] = None, nthreads: int = 1, corpus_is_bf16: bool = False,
                fast: bool = False):
    """oracle.c seq scan: float accumulators, heap top-k.  corpus fp32 [N,D] or uint16 bf16 bits.
    ``fast`` = pgvector's own SIMD-reassociating build flags (timing only)."""
    lib = clib(fast)
    corpus = np.ascontiguousarray(corpus, dtype=np.uint16 if corpus_is_bf16 else np.float32)
    q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
    n, d = corpus.shape
    nq = q.shape[0]
    out_d = np.empty((nq, k), dtype=np.float64)
    out_i = np.empty((nq, k), dtype=np.int64)
    m = pack_mask(mask)
    lib.orc_scan_topk_batch(METRICS[metric], _ptr(corpus), int(corpus_is_bf16), n, d, _ptr(q), nq, k,
                            _ptr(m), _ptr(out_d), _ptr(out_i), nthreads)
    return out_d, out_i


def c_hybrid_topk(metric: str, corpus: np.ndarray, query: np.ndarray, bm25: Optional[np.ndarray],
                  semantic_weight: float, bm25_weight: float, k: int,
                  mask: Optional[np.ndarray] = None, corpus_is_bf16: bool = False):
    lib = clib()
    corpus = np.ascontiguousarray(corpus, dtype=np.uint16 if corpus_is_bf16 else np.float32)
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
    n, d = corpus.shape
    out_c = np.empty(k, dtype=np.float64)
    out_i = np.empty(k, dtype=np.int64)
    b = None if bm25 is None else np.ascontiguousarray(bm25, dtype=np.float64)
    m = pack_mask(mask)
    cnt = lib.orc_hybrid_topk(METRICS[metric], _ptr(corpus), int(corpus_is_bf16), n, d, _ptr(q), _ptr(b),
                              float(semantic_weight), float(bm25_weight), k, _ptr(m), _ptr(out_c), _ptr(out_i))
    return out_c[:cnt], out_i[:cnt]


def c_pool_normalize(hidden: np.ndarray, mask: np.ndarray) -> np.ndarray:
    lib = clib()
    h = np.ascontiguousarray(hidden, dtype=np.float32)
    m = np.ascontiguousarray(mask, dtype=np.int64)
    B, L, H = h.shape
    out = np.empty((B, H), dtype=np.float32)
    lib.orc_pool_normalize(_ptr(h), _ptr(m), B, L, H, _ptr(out))
    return out


# --------------------------------------------------------------------------------------------
# hnsw.c via ctypes: restated default index of the reference store (recall report only)
# --------------------------------------------------------------------------------------------
_HNSW = None


def hnsw_lib() -> ctypes.CDLL:
    global _HNSW
    if _HNSW is None:
        so = os.path.join(_HERE, f"libhnsw_{_cpu_tag()}.so")
        src = os.path.join(_HERE, "hnsw.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O3", "-march=native", "-ffast-math", "-fPIC", "-shared", "-o", so, src, "-lm"])
        lib = ctypes.CDLL(so)
        lib.hnsw_create.restype = ctypes.c_void_p
        lib.hnsw_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        lib.hnsw_destroy.argtypes = [ctypes.c_void_p]
        lib.hnsw_insert.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.hnsw_search.restype = ctypes.c_int
        lib.hnsw_search.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _HNSW = lib
    return _HNSW


class HnswIndex:
    """The reference's default index, restated: HNSW(m=16, ef_construction=64) under cosine, queried
    with pgvector's default hnsw.ef_search = 40 [external] (init.sql:280-284)."""

    def __init__(self, unit_rows: np.ndarray, m: int = 16, ef_construction: int = 64, seed: int = 1):
        self.lib = hnsw_lib()
        self.data = np.ascontiguousarray(unit_rows, dtype=np.float32)
        n, d = self.data.shape
        self.h = self.lib.hnsw_create(_ptr(self.data), n, d, m, ef_construction, seed)
        for i in range(n):
            self.lib.hnsw_insert(self.h, i)

    def search(self, queries: np.ndarray, k: int, ef_search: int = 40):
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        q = q / np.linalg.norm(q, axis=1, keepdims=True)
        ids = np.empty((q.shape[0], k), dtype=np.int32)
        dist = np.empty((q.shape[0], k), dtype=np.float32)
        for i in range(q.shape[0]):
            self.lib.hnsw_search(self.h, _ptr(q[i]), k, max(ef_search, 1), _ptr(ids[i]), _ptr(dist[i]))
        # pgvector returns at most ef_search rows from an index scan [external]
        if k > ef_search:
            ids[:, ef_search:] = -1
        return dist, ids.astype(np.int64)

    def close(self):
        if self.h:
            self.lib.hnsw_destroy(self.h)
            self.h = None


def recall_at_k(ids_approx: np.ndarray, ids_exact: np.ndarray) -> float:
    hits = sum(len(set(a[a >= 0].tolist()) & set(e.tolist())) for a, e in zip(ids_approx, ids_exact))
    return hits / float(ids_exact.size)
