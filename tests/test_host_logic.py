"""CPU tests: the C-ABI library loads and exports what include/archi_b200.h declares, the host-side
mirror honours the reference's constructor / retriever contracts, and the multi-GPU host logic works
over gloo with world_size 2.  No GPU compute is called here."""
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the boundary -------------------------------------------------------------------------------------------
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "archi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(archi_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from archi_b200 import _native as N
    from archi_b200.build import build
    build()
    L = N.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in archi_b200.h but not exported"
    assert sorted(N.EXPORTS) == declared
    assert L.archi_abi_version() == 2


def test_no_silent_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from archi_b200.store import NativeStore
    with pytest.raises(RuntimeError, match="libarchi_b200"):
        NativeStore(384)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "archi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("the oracle uses", ""), f"{f} mentions the oracle"


# ---- constructor contract (tests/unit/test_postgres_vectorstore.py:87-137) -------------------------------------
def test_vectorstore_init_contract():
    from archi_b200 import B200VectorStore
    s = B200VectorStore(pg_config={"host": "x"}, embedding_function="emb", collection_name="t_init_default")
    assert s._distance_metric == "cosine" and s._distance_op == "<=>" and s.embeddings == "emb"
    assert B200VectorStore(None, None, "t_init_l2", "l2")._distance_op == "<->"
    assert B200VectorStore(None, None, "t_init_ip", "inner_product")._distance_op == "<#>"
    with pytest.raises(ValueError, match="distance_metric must be one of"):
        B200VectorStore(None, None, "t_init_bad", "invalid")
    # an empty collection answers without touching the GPU
    assert s.count() == 0
    assert s.similarity_search_by_vector([0.0] * 4, k=3) == []
    assert s.add_texts([]) == []
    assert s.delete() is False


def test_json_text_filter_semantics():
    from archi_b200.vectorstore import _json_text
    assert _json_text(5) == "5" and _json_text("a") == "a" and _json_text(True) == "true"


# ---- retriever policies (hybrid_retriever.py:64-105) ----------------------------------------------------------------
class _FakeStore:
    def __init__(self, exc=None):
        self.exc, self.calls = exc, []

    def hybrid_search(self, query, k, semantic_weight, bm25_weight):
        self.calls.append(("hybrid", k, semantic_weight, bm25_weight))
        if self.exc:
            raise self.exc
        return [("doc", 0.865)]

    def similarity_search_with_score(self, query, k):
        self.calls.append(("semantic", k))
        return [("doc", 0.9)]

    def similarity_search(self, query, k):
        self.calls.append(("plain", k))
        return ["doc"]


def test_hybrid_retriever_defaults_and_delegation():
    from archi_b200 import HybridRetriever
    vs = _FakeStore()
    r = HybridRetriever(vs)
    assert (r.k, r.bm25_weight, r.semantic_weight) == (5, 0.5, 0.5)
    assert r.invoke("q") == [("doc", 0.865)]
    assert vs.calls == [("hybrid", 5, 0.5, 0.5)]
    r = HybridRetriever(_FakeStore(), k=7, bm25_weight=0.6, semantic_weight=0.4)
    r.invoke("q")
    assert r.vectorstore.calls == [("hybrid", 7, 0.4, 0.6)]


def test_hybrid_retriever_error_policy():
    from archi_b200 import HybridRetriever
    vs = _FakeStore(RuntimeError("hybrid search is not supported by this backend"))
    assert HybridRetriever(vs, k=2).invoke("q") == [("doc", 0.9)]
    assert vs.calls[-1] == ("semantic", 2)
    vs = _FakeStore(RuntimeError("Hybrid search requires pg_textsearch BM25 index on document_chunks; none found."))
    with pytest.raises(RuntimeError, match="BM25 index"):
        HybridRetriever(vs).invoke("q")

    class NoHybrid:
        def similarity_search_with_score(self, query, k):
            return [("d", 1.0)]
    assert HybridRetriever(NoHybrid(), k=1).invoke("q") == [("d", 1.0)]


def test_semantic_and_grading_retrievers():
    from archi_b200 import GradingRetriever, SemanticRetriever
    from archi_b200.retrievers import make_instruction_query
    cfg = {"embedding_name": "HF", "embedding_class_map": {"HF": {"kwargs": {"model_name": "Qwen/Qwen3-Embedding-0.6B"}}}}
    seen = []

    class VS(_FakeStore):
        def similarity_search_with_score(self, query, k):
            seen.append(query)
            return super().similarity_search_with_score(query, k)
    r = SemanticRetriever(VS(), cfg, instructions="find docs")
    assert r.k == 3 and r.invoke("what?") == [("doc", 0.9)]
    assert seen[-1] == make_instruction_query("find docs", "what?") == "Instruct: find docs\nQuery:what?"
    cfg2 = {"embedding_name": "HF", "embedding_class_map": {"HF": {"kwargs": {"model_name": "all-MiniLM-L6-v2"}}}}
    SemanticRetriever(VS(), cfg2, instructions="find docs").invoke("what?")
    assert seen[-1] == "what?"
    g = GradingRetriever(_FakeStore())
    assert g.k == 3 and g.invoke("q") == ["doc"]


# ---- row sharding ---------------------------------------------------------------------------------------------------
def test_plan_row_shards():
    from archi_b200.sharded import offsets_from_counts, plan_row_shards
    assert plan_row_shards(10, 4) == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert plan_row_shards(0, 2) == [(0, 0), (0, 0)]
    assert plan_row_shards(3, 8)[3:] == [(3, 0)] * 5
    p = plan_row_shards(100_000_000, 8)
    assert sum(c for _, c in p) == 100_000_000 and p[0] == (0, 12_500_000)
    assert offsets_from_counts([3, 0, 5]) == [0, 3, 3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, metric, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from archi_b200.sharded import ShardedStore, plan_row_shards
    from oracle import oracle as orc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(99)
    corpus = rng.standard_normal((1001, 32)).astype(np.float32)
    queries = rng.standard_normal((6, 32)).astype(np.float32)
    k = 7
    first, cnt = plan_row_shards(corpus.shape[0], world)[rank]
    shard = corpus[first:first + cnt]
    larger = metric == "cosine"

    def local_search(q, kk, id_offset):     # stand-in for the CUDA scan: the oracle on this shard
        d, i = orc.exact_topk(metric, shard, q.numpy(), kk)
        s = orc.score_from_distance(metric, d)
        pad = kk - s.shape[1]
        if pad:
            s = np.concatenate([s, np.full((s.shape[0], pad), np.nan)], 1)
            i = np.concatenate([i, np.full((i.shape[0], pad), -1)], 1)
        i = np.where(i >= 0, i + id_offset, -1)
        return torch.from_numpy(s.astype(np.float32)), torch.from_numpy(i)

    def merge(scores, ids, larger_is_better):  # stand-in for archi_merge_topk
        s = scores.numpy().transpose(1, 0, 2).reshape(scores.shape[1], -1)
        i = ids.numpy().transpose(1, 0, 2).reshape(ids.shape[1], -1)
        key = np.where(i >= 0, s if larger_is_better else -s, -np.inf)
        order = np.lexsort((i, -key), axis=1)[:, :scores.shape[2]]
        return torch.from_numpy(np.take_along_axis(s, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))

    st = ShardedStore(None, larger_is_better=larger, local_search=local_search, merge=merge, local_rows=lambda: cnt)
    st.sync_layout()
    assert st.id_offset == first and st.total_rows == corpus.shape[0]
    s, i = st.search(torch.from_numpy(queries), k)
    d_true, i_true = orc.exact_topk(metric, corpus, queries, k)
    ok = all(orc.same_topk_up_to_ties(i[q].tolist(), i_true[q], d_true[q], rel_tol=1e-6) for q in range(6))
    ok = ok and np.allclose(s.numpy(), orc.score_from_distance(metric, d_true), rtol=1e-5, atol=1e-6)
    open(os.path.join(tmp, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    dist.destroy_process_group()


@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_sharded_search_gloo_world2(tmp_path, metric):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, metric, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm: oracle.c on the host cores) keeps stdout to exactly one
    JSON line carrying the contract's keys; nothing of archi_b200's CUDA path is needed for it."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "queries_per_sec_exact_top10" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["rows"] == 1_000_000 and d["config"]["dim"] == 384 and d["config"]["k"] == 10


def test_probe_threshold_invariant():
    """What the tensor path's probe launch relies on (csrc/tensor.cu, tc_maxima_threshold_kernel): the
    k'-th largest of the per-32-column chunk maxima of ANY subset of the rows is a lower bound of the
    k'-th largest key over all rows, because chunk maxima belong to distinct rows -- so no row of the
    final top-k' can be rejected by it.  Masked rows (key -inf) only lower the bound."""
    rng = np.random.default_rng(17)
    for trial in range(20):
        n, kprime = int(rng.integers(2000, 20000)), int(rng.choice([32, 64, 224]))
        keys = rng.standard_normal(n).astype(np.float32)
        if trial % 3 == 0:
            keys[rng.random(n) < 0.3] = -np.inf                       # tombstones / filter
        if trial % 4 == 0:
            keys = np.round(keys, 1)                                   # heavy ties
        probe = keys[: (n // 12) // 32 * 32].reshape(-1, 32)           # ~1/12 of the rows, whole chunks
        maxima = probe.max(axis=1)
        valid = maxima[maxima > -np.inf]
        kth_all = np.sort(keys)[::-1][kprime - 1]
        if valid.size >= kprime:
            thr = np.sort(valid)[::-1][kprime - 1]
            assert thr <= kth_all
            # rows strictly above the threshold are what the main scan admits: at least ... the top ones
            assert (keys > thr).sum() >= min(kprime - 1, (keys > kth_all).sum())


def test_token_budget_batching_of_b200_embeddings():
    """B200Embeddings._batches (host logic, no GPU): consecutive texts share a forward until sequences x padded length
    would pass max_batch_tokens; lengths are padded to a multiple of 32 (capped at max_seq_length); order, ids and
    masks are preserved; an explicit batch_size restores fixed-count batches."""
    import random
    from archi_b200 import embeddings as E

    class Stub(E.B200Embeddings):
        def __init__(self, batch_size=None, budget=1024, max_len=64):      # no model, no device
            self.batch_size, self.max_batch_tokens, self.max_seq_length = batch_size, budget, max_len
            self.tokenizer = E.HashTokenizer()
            self.calls = []

        def _forward_ids(self, ids, mask):
            self.calls.append((ids.copy(), mask.copy()))
            return ids, mask

        def _forward(self, texts):
            return self._forward_ids(*self.tokenizer(texts, self.max_seq_length))

    rng = random.Random(1)
    texts = [" ".join("w%d" % rng.randrange(1000) for _ in range(rng.randint(1, 100))) for _ in range(203)]
    texts.sort(key=len)                      # what IngestionDriver hands over
    st = Stub()
    list(st._batches(texts))
    shapes = [c[0].shape for c in st.calls]
    assert sum(s[0] for s in shapes) == len(texts)
    assert all(s[1] % 32 == 0 and s[1] <= 64 for s in shapes)
    assert all(s[0] * s[1] <= 1024 or s[0] == 1 for s in shapes)
    assert len(shapes) < len(texts) / 8                                     # few, large forwards
    ids_all, mask_all = st.tokenizer(texts, 64)
    at = 0
    for ids, mask in st.calls:
        b, w = ids.shape[0], min(ids.shape[1], ids_all.shape[1])
        assert np.array_equal(ids[:, :w], ids_all[at:at + b, :w]) and np.array_equal(mask[:, :w], mask_all[at:at + b, :w])
        assert (mask[:, w:] == 0).all() and (mask.sum(axis=1) >= 2).all()
        at += b
    # a single very long text still gets its own forward, truncated to max_seq_length
    st1 = Stub(budget=16)
    list(st1._batches(["a " * 500, "b"]))
    assert [c[0].shape for c in st1.calls] == [(1, 64), (1, 32)]
    # explicit batch size: fixed-count batches padded to the longest member
    st2 = Stub(batch_size=32)
    list(st2._batches(texts))
    assert [c[0].shape[0] for c in st2.calls] == [32] * 6 + [11]
