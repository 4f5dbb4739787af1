"""Generates tests/golden/*.npz: seeded inputs + the oracle's outputs for them.

The reference's arithmetic cannot run here (pgvector / pg_textsearch / sentence-transformers are
absent) and its tests hold no numeric golden vectors for this path, so these
fixtures pin the ORACLE (fp64 numpy truth, cross-checked against oracle.c at generation time)
rather than the reference -- "parity unpinned".  They guard against silent drift of the oracle and
give the GPU tests a committed input/output set.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def main():
    rng = np.random.default_rng(20261017)
    # -- search: 2048 x 96 fp32 corpus (not unit norm for l2/ip variety), 5 queries, k = 10 ----
    corpus = (unit_rows(rng, 2048, 96) * rng.uniform(0.5, 2.0, size=(2048, 1))).astype(np.float32)
    corpus[7] = corpus[3]                    # an exact duplicate row: a tie on every metric
    queries = unit_rows(rng, 5, 96)
    queries[4] = corpus[100] / np.linalg.norm(corpus[100])
    out = {"corpus": corpus, "queries": queries}
    for metric in orc.METRICS:
        d, i = orc.exact_topk(metric, corpus, queries, 10)
        dc, ic = orc.c_scan_topk(metric, corpus, queries, 10)
        assert (i == ic).all(), metric
        assert np.allclose(d, dc, rtol=2e-6, atol=2e-6), metric
        out[f"{metric}_dist"] = d
        out[f"{metric}_ids"] = i
    # bf16 storage: the stored (rounded) values are what is searched
    bits = orc.f32_to_bf16_bits(corpus)
    stored = orc.bf16_bits_to_f32(bits)
    d, i = orc.exact_topk("cosine", stored, queries, 10)
    out["bf16_bits"] = bits
    out["bf16_cosine_dist"] = d
    out["bf16_cosine_ids"] = i
    np.savez_compressed(os.path.join(HERE, "search_2048x96.npz"), **out)

    # -- hybrid: 512 x 64, sparse bm25 with NULLs -------------------------------------------------
    c2 = unit_rows(rng, 512, 64)
    q2 = unit_rows(rng, 1, 64)[0]
    bm25 = np.full(512, np.nan)
    hit = rng.choice(512, size=40, replace=False)
    bm25[hit] = rng.uniform(0.1, 6.0, size=40)
    h = {"corpus": c2, "query": q2, "bm25": bm25}
    for metric in orc.METRICS:
        for ws, wb in ((0.7, 0.3), (0.4, 0.6)):
            comb, ids = orc.exact_hybrid_topk(metric, c2, q2, bm25, ws, wb, 8)
            cc, ic = orc.c_hybrid_topk(metric, c2, q2, bm25, ws, wb, 8)
            assert (ids == ic).all()
            h[f"{metric}_{ws}_{wb}_combined"] = comb
            h[f"{metric}_{ws}_{wb}_ids"] = ids
    np.savez_compressed(os.path.join(HERE, "hybrid_512x64.npz"), **h)

    # -- pool + normalise: B=6, L=24, H=64 with ragged masks incl. an all-zero mask --------------------
    hidden = rng.standard_normal((6, 24, 64)).astype(np.float32)
    lens = [24, 1, 7, 16, 0, 23]
    mask = np.zeros((6, 24), dtype=np.int64)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    pooled = orc.pool_normalize(hidden, mask)
    pc = orc.c_pool_normalize(hidden, mask)
    assert np.allclose(pooled, pc, atol=1e-6)
    np.savez_compressed(os.path.join(HERE, "pool_6x24x64.npz"), hidden=hidden, mask=mask, pooled=pooled)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
