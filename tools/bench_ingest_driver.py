"""End-to-end ingest through IngestionDriver (SURVEY.md 8f-1, config 5's chunks/s): synthetic text files
-> load + split (thread pool) -> cross-file, length-ordered embedding (B200Embeddings: encoder forward +
fused pool/normalise) -> add_embedded_texts (append kernel + lexical index) -> statuses and commits.

    python tools/bench_ingest_driver.py [--files 400] [--paragraphs 40] [--per-file]   # on a B200
    python tools/bench_ingest_driver.py --dry-run                                      # CPU: host logic only

--per-file embeds one file per call (the reference's loop, manager.py:362-373) for comparison.
Prints one JSON line."""
import argparse
import json
import os
import random
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from archi_b200 import B200VectorStore, IngestionDriver


def write_corpus(root, n_files, n_paragraphs, seed):
    rng = random.Random(seed)
    words = ["w%d" % i for i in range(30000)]
    files = {}
    for i in range(n_files):
        paras = [" ".join(rng.choice(words) for _ in range(rng.randint(20, 160))) for _ in range(n_paragraphs)]
        path = os.path.join(root, f"doc{i:05d}.md")
        with open(path, "w") as f:
            f.write("\n\n".join(paras))
        files[f"hash{i:05d}"] = path
    return files


class DryEmbeddings:
    """CPU stand-in (--dry-run): unit-norm pseudo-random rows, one per text."""

    def __init__(self, dim=384):
        self.dim = dim

    def embed_documents(self, texts):
        rng = np.random.default_rng(len(texts))
        x = rng.standard_normal((len(texts), self.dim)).astype(np.float32)
        return x / np.linalg.norm(x, axis=1, keepdims=True)

    def embed_query(self, text):
        return self.embed_documents([text])[0]


class DryNative:
    def __init__(self, dim):
        self.dim, self.n = dim, 0

    def append(self, emb):
        first = self.n
        self.n += len(emb)
        return first

    def close(self):
        pass


def run(n_files=400, n_paragraphs=40, chunk_size=1000, per_file=False, bm25=True, dry_run=False, ef=None, device=0):
    """One timed IngestionDriver.add_files over a synthetic corpus; returns the record."""
    with tempfile.TemporaryDirectory() as root:
        files = write_corpus(root, n_files, n_paragraphs, seed=5)
        if dry_run:
            ef = DryEmbeddings()
        elif ef is None:
            import torch
            from archi_b200 import B200Embeddings
            assert torch.cuda.is_available(), "needs a GPU (or --dry-run)"
            ef = B200Embeddings(device=device)
        kw = {} if dry_run else {"device": device}
        B200VectorStore.drop_collection("ingest_bench", **kw)
        store = B200VectorStore({}, ef, collection_name="ingest_bench", storage_dtype="f32" if dry_run else "bf16",
                                bm25_index=bm25, **kw)
        if dry_run:
            fake = DryNative(ef.dim)
            store._coll.ensure_native = lambda dim, shard=0: fake
        driver = IngestionDriver(store, chunk_size=chunk_size, commit_batch_size=1 if per_file else 25)
        if not dry_run:                       # warm-up: CUDA context, encoder autotuning, lazy buffers
            import torch
            warm = dict(list(files.items())[:8])
            driver.add_files(warm)
            torch.cuda.synchronize()
            B200VectorStore.drop_collection("ingest_bench", **kw)
            store = B200VectorStore({}, ef, collection_name="ingest_bench", storage_dtype="bf16", bm25_index=bm25, **kw)
            driver = IngestionDriver(store, chunk_size=chunk_size, commit_batch_size=1 if per_file else 25)
        t0 = time.perf_counter()
        report = driver.add_files(files)
        if not dry_run:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out = {"metric": "ingest_chunks_per_sec", "value": report.chunks / dt, "unit": "chunks/s", "files": n_files,
               "chunks": report.chunks, "seconds": dt, "embed_calls": report.embed_calls, "commits": report.commits,
               "failed": len(report.failed), "mode": "per-file" if per_file else "cross-file groups of 25",
               "bm25_index": bm25, "dry_run": dry_run, "embeddings": type(ef).__name__}
        B200VectorStore.drop_collection("ingest_bench", **kw)
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=400)
    ap.add_argument("--paragraphs", type=int, default=40)
    ap.add_argument("--chunk-size", type=int, default=1000)
    ap.add_argument("--per-file", action="store_true", help="commit_batch_size=1: one embedding call per file")
    ap.add_argument("--no-bm25", action="store_true")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()
    print(json.dumps(run(args.files, args.paragraphs, args.chunk_size, args.per_file, not args.no_bm25, args.dry_run)))


if __name__ == "__main__":
    main()
