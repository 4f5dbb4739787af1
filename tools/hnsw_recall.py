"""Recall@k of the reference store's DEFAULT index (restated HNSW m=16, ef_construction=64, ef_search=40,
cosine) against the exact result, on synthetic unit-norm embeddings.  CPU only.

    python tools/hnsw_recall.py [--rows 100000] [--dim 384] [--queries 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--queries", type=int, default=200)
args = ap.parse_args()
rng = np.random.default_rng(1234)


def unit(a):
    return (a / np.linalg.norm(a, axis=1, keepdims=True)).astype(np.float32)


def report(x, q, label):
    t0 = time.perf_counter()
    idx = orc.HnswIndex(x)
    res = {"data": label, "build_s": time.perf_counter() - t0}
    for k in (5, 10, 100):
        _, exact = orc.exact_topk("cosine", x, q, k)
        t0 = time.perf_counter()
        _, approx = idx.search(q, k, ef_search=40)
        res[f"recall@{k}_ef_search_40"] = orc.recall_at_k(approx, exact)
        res[f"search_ms_per_query_k{k}"] = (time.perf_counter() - t0) * 1e3 / q.shape[0]
    _, exact = orc.exact_topk("cosine", x, q, 10)
    for ef in (10, 100, 400):
        _, approx = idx.search(q, 10, ef_search=ef)
        res[f"recall@10_ef_search_{ef}"] = orc.recall_at_k(approx, exact)
    idx.close()
    return res


out = {"index": "HNSW m=16 ef_construction=64 (restated; reference default, init.sql:280-284), cosine, "
                "pgvector default hnsw.ef_search=40 [external]; an index scan returns at most ef_search rows",
       "rows": args.rows, "dim": args.dim, "queries": args.queries,
       "exact_store_recall": 1.0, "datasets": []}
# (1) the benchmark's own distribution: isotropic gaussian rows -- the worst case for a graph index
out["datasets"].append(report(unit(rng.standard_normal((args.rows, args.dim))), unit(rng.standard_normal((args.queries, args.dim))),
                              "isotropic unit-norm gaussian (the bench distribution; intrinsic dimension = dim)"))
# (2) embedding-like rows: a 12-d latent mixed into dim dimensions plus 5 % noise
A = rng.standard_normal((12, args.dim))
xl = rng.standard_normal((args.rows, 12)) @ A + 0.05 * rng.standard_normal((args.rows, args.dim))
ql = rng.standard_normal((args.queries, 12)) @ A + 0.05 * rng.standard_normal((args.queries, args.dim))
out["datasets"].append(report(unit(xl), unit(ql), "12-d latent embedded in dim dimensions + 5% noise (embedding-like)"))
print(json.dumps(out))
