"""BASELINE config 5: hybrid BM25 + dense retrieval on 1M synthetic chunks (fused score fusion) and
end-to-end ingest (encoder forward + fused pool/normalise/append) throughput in chunks/sec.

    python tools/bench_config5.py [--rows 1000000] [--queries 200] [--ingest-batches 8]

Prints one JSON object.  Synthetic data: unit-norm fp32 chunk embeddings (D = 384), documents of 24
Zipf-distributed term ids over a 50k vocabulary, 3-term queries; the encoder is a random-init
MiniLM-L6-shaped BertModel in bf16 (no weights offline), sequences of 256 tokens.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--queries", type=int, default=200)
    ap.add_argument("--ingest-batches", type=int, default=8)
    ap.add_argument("--ingest-batch", type=int, default=1024)
    ap.add_argument("--seq-len", type=int, default=256)
    args = ap.parse_args()

    import torch
    from archi_b200.bm25 import LexicalIndex
    from archi_b200.store import NativeStore, pool_normalize

    dev = torch.device("cuda", 0)
    peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    out = {"config": "configs[4]: hybrid BM25 + dense on synthetic chunks, plus ingest", "rows": args.rows, "dim": args.dim}

    # ---------------- hybrid -----------------------------------------------------------------------
    g = torch.Generator(device=dev).manual_seed(1234 + 5000)
    store = NativeStore(args.dim, "cosine", "f32", capacity_rows=args.rows)
    for s in range(0, args.rows, 262144):
        m = min(262144, args.rows - s)
        x = torch.randn((m, args.dim), generator=g, device=dev)
        store.append(x / x.norm(dim=1, keepdim=True))
    rng = np.random.default_rng(5)
    vocab, doc_len = 50_000, 24
    t0 = time.perf_counter()
    lex = LexicalIndex(device=0)
    tokens = (rng.zipf(1.3, size=(args.rows, doc_len)) % vocab).astype(np.int64)
    lex.add_token_matrix(tokens)
    q_terms = [(rng.zipf(1.3, size=3) % vocab).astype(np.int64) for _ in range(args.queries)]
    lex.score(q_terms[0])                           # builds the device posting lists
    torch.cuda.synchronize()
    out["bm25_index_build_s"] = time.perf_counter() - t0
    q = torch.randn((args.queries, args.dim), generator=g, device=dev)
    q = q / q.norm(dim=1, keepdim=True)
    bm = torch.zeros(args.rows, dtype=torch.float32, device=dev)

    def hybrid_one(i):
        lex.score(q_terms[i], out=bm)
        return store.search(q[i:i + 1], 5, bm25=bm[None, :], semantic_weight=0.4, bm25_weight=0.6, hybrid=True)

    def dense_one(i):
        return store.search(q[i:i + 1], 5)

    for fn, name in ((hybrid_one, "hybrid"), (dense_one, "dense_only")):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.queries):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.queries
        out[name] = {"queries_per_s": 1e3 / ms, "ms_per_query": ms, "k": 5, "batch": 1}
    nnz = [int((lex.score(t) != 0).sum().item()) for t in q_terms[:20]]
    out["hybrid"]["rows_with_bm25_match_mean"] = float(np.mean(nnz))
    out["hybrid"]["algorithmic_bytes_per_query"] = args.rows * args.dim * 4 + args.rows * 4
    out["hybrid"]["achieved_GBps_whole_query"] = out["hybrid"]["algorithmic_bytes_per_query"] / (out["hybrid"]["ms_per_query"] * 1e-3) / 1e9
    out["hybrid"]["frac_hbm_whole_query"] = out["hybrid"]["achieved_GBps_whole_query"] / hbm_peak
    store.close()

    # ---------------- ingest ------------------------------------------------------------------------
    from transformers import BertConfig, BertModel
    from archi_b200.embeddings import MINILM_L6
    torch.manual_seed(0)
    model = BertModel(BertConfig(**MINILM_L6), add_pooling_layer=False).to(dev, torch.bfloat16).eval()
    B, L, H = args.ingest_batch, args.seq_len, MINILM_L6["hidden_size"]
    ids = torch.randint(1000, 30000, (B, L), device=dev)
    lens = torch.randint(L // 2, L + 1, (B,), device=dev)
    mask = (torch.arange(L, device=dev)[None, :] < lens[:, None]).to(torch.int64)
    sink = NativeStore(H, "cosine", "bf16", capacity_rows=B * (args.ingest_batches + 4))

    def step():
        with torch.inference_mode():
            hidden = model(input_ids=ids, attention_mask=mask).last_hidden_state
        sink.pool_normalize_append(hidden, mask)
        return hidden

    for _ in range(2):
        hidden = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.ingest_batches):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.ingest_batches
    # the fused kernel alone
    for _ in range(3):
        pool_normalize(hidden, mask, want_bf16=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        pool_normalize(hidden, mask, want_bf16=True)
    e1.record()
    torch.cuda.synchronize()
    ms_pool = e0.elapsed_time(e1) / 20
    live_tokens = int(mask.sum().item())
    pool_bytes = live_tokens * H * 2 + B * L * 8 + B * H * (4 + 2)   # masked tokens are not read
    out["ingest"] = {"chunks_per_s": B / (ms_step * 1e-3), "ms_per_batch": ms_step, "batch": B, "seq_len": L,
                     "encoder": "BertModel MiniLM-L6 shape, random init, bf16 (PyTorch)",
                     "pool_normalize_ms": ms_pool, "pool_algorithmic_bytes": pool_bytes,
                     "pool_achieved_GBps": pool_bytes / (ms_pool * 1e-3) / 1e9,
                     "pool_frac_hbm": pool_bytes / (ms_pool * 1e-3) / 1e9 / hbm_peak,
                     "pool_share_of_step": ms_pool / ms_step}
    sink.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
