"""GPU tests of the pool+normalise kernel and of the reference-facing surface (B200VectorStore,
retrievers, embeddings) -- the behavioural counterparts of the reference's mocked unit tests
(tests/unit/test_postgres_vectorstore.py), checked against the oracle instead of a mocked cursor."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


# ---- pool + normalise ------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("shape", [(6, 24, 64), (3, 256, 384), (33, 17, 768), (2, 5, 1024), (4, 9, 8)])
@pytest.mark.parametrize("mask_dtype", ["i64", "i32"])
@pytest.mark.parametrize("ring", ["0", "1"])
def test_pool_normalize_matches_oracle(dtype, shape, mask_dtype, ring, monkeypatch):
    """ring=1 forces the cp.async.bulk ring kernel (persistent CTAs) on these small batches, ring=0 the
    one-CTA-per-sequence kernel; both must agree with the oracle."""
    import torch
    from archi_b200.store import pool_normalize
    monkeypatch.setenv("ARCHI_POOL_RING", ring)
    B, L, H = shape
    rng = np.random.default_rng(B * 1000 + L)
    hidden = rng.standard_normal(shape).astype(np.float32)
    lens = rng.integers(1, L + 1, size=B)
    lens[0] = L
    mask = (np.arange(L)[None, :] < lens[:, None]).astype(np.int64)
    if B > 2:
        mask[2, ::2] = 0                           # holes inside the sequence, not just padding
        mask[2, 1] = 1
    h_t = torch.from_numpy(hidden).cuda()
    if dtype == "bf16":
        h_t = h_t.to(torch.bfloat16)
        hidden = h_t.float().cpu().numpy()          # the oracle sees the same (rounded) inputs
    m_t = torch.from_numpy(mask).cuda().to(torch.int64 if mask_dtype == "i64" else torch.int32)
    out_f32, out_bf16 = pool_normalize(h_t, m_t, want_bf16=True)
    want = orc.pool_normalize(hidden, mask)
    got = out_f32.cpu().numpy()
    assert np.allclose(got, want, rtol=1e-5, atol=2e-6)
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    # the bf16 copy is the round-to-nearest-even cast of the fp32 result
    want_bits = orc.f32_to_bf16_bits(got)
    got_bits = out_bf16.view(torch.int16).cpu().numpy().view(np.uint16)
    assert np.array_equal(got_bits, want_bits)


@pytest.mark.parametrize("dtype,shape", [("bf16", (1024, 256, 384)), ("bf16", (700, 64, 768)), ("f32", (600, 40, 1024)),
                                         ("f32", (2100, 9, 384))])
def test_pool_normalize_machine_filling_batches(dtype, shape):
    """Batches of >= 2 sequences per SM take the ring kernel by default: several sequences per persistent CTA,
    the ring running ahead into the next sequence; all-masked rows, holes, full and one-token sequences included.
    The one-CTA-per-sequence kernel must give the same rows up to summation order."""
    import torch
    from archi_b200.store import pool_normalize
    B, L, H = shape
    rng = np.random.default_rng(B + L)
    hidden = rng.standard_normal(shape, dtype=np.float32)
    lens = rng.integers(1, L + 1, size=B)
    lens[:4] = [L, 1, L, 2]
    mask = (np.arange(L)[None, :] < lens[:, None]).astype(np.int64)
    mask[5, :] = 0
    mask[6, ::3] = 0
    mask[B - 1, :] = 0
    h_t = torch.from_numpy(hidden).cuda()
    if dtype == "bf16":
        h_t = h_t.to(torch.bfloat16)
        hidden = h_t.float().cpu().numpy()
    m_t = torch.from_numpy(mask).cuda()
    out_f32, out_bf16 = pool_normalize(h_t, m_t, want_bf16=True)
    got = out_f32.cpu().numpy()
    want = orc.pool_normalize(hidden, mask)
    assert np.allclose(got, want, rtol=1e-5, atol=2e-6)
    assert np.array_equal(got[5], np.zeros(H, dtype=np.float32)) and np.array_equal(got[B - 1], np.zeros(H, dtype=np.float32))
    assert np.array_equal(out_bf16.view(torch.int16).cpu().numpy().view(np.uint16), orc.f32_to_bf16_bits(got))
    os.environ["ARCHI_POOL_RING"] = "0"
    try:
        other, _ = pool_normalize(h_t, m_t)
    finally:
        del os.environ["ARCHI_POOL_RING"]
    assert np.allclose(other.cpu().numpy(), got, rtol=1e-5, atol=2e-6)
    # repeated launches are bit-identical (fixed summation order)
    again, _ = pool_normalize(h_t, m_t)
    assert torch.equal(again, out_f32)


@pytest.mark.parametrize("ring", ["0", "1"])
def test_pool_normalize_golden_and_all_zero_mask(golden_dir, ring, monkeypatch):
    import torch
    from archi_b200.store import pool_normalize
    monkeypatch.setenv("ARCHI_POOL_RING", ring)
    p = np.load(os.path.join(golden_dir, "pool_6x24x64.npz"))
    out, _ = pool_normalize(torch.from_numpy(p["hidden"]).cuda(), torch.from_numpy(p["mask"]).cuda())
    got = out.cpu().numpy()
    assert np.allclose(got, p["pooled"], rtol=1e-5, atol=2e-6)
    assert np.array_equal(got[4], np.zeros(64, dtype=np.float32))      # mask all zero -> zero row, no NaN


@pytest.mark.parametrize("storage", ["f32", "bf16"])
@pytest.mark.parametrize("ring", ["0", "1"])
def test_pool_normalize_append_equals_pool_then_append(storage, ring, monkeypatch):
    import torch
    from archi_b200.store import NativeStore, pool_normalize
    monkeypatch.setenv("ARCHI_POOL_RING", ring)
    rng = np.random.default_rng(4)
    hidden = torch.from_numpy(rng.standard_normal((10, 12, 96)).astype(np.float32)).cuda()
    mask = torch.ones((10, 12), dtype=torch.int64, device="cuda")
    mask[3, 5:] = 0
    a = NativeStore(96, "cosine", storage)
    a.append(torch.zeros((3, 96), device="cuda"))                     # rows already present
    assert a.pool_normalize_append(hidden, mask) == 3
    pooled, _ = pool_normalize(hidden, mask)
    b = NativeStore(96, "cosine", storage)
    b.append(torch.zeros((3, 96), device="cuda"))
    b.append(pooled)
    assert a.rows() == b.rows() == 13 and a.count() == 13
    assert np.array_equal(a.read_rows(0, 13), b.read_rows(0, 13))
    q = rng.standard_normal((2, 96)).astype(np.float32)
    ra, rb = a.search(q, 5), b.search(q, 5)
    assert np.array_equal(ra[1], rb[1]) and np.allclose(ra[0], rb[0], rtol=1e-6, atol=1e-7)
    a.close()
    b.close()


# ---- the vectorstore surface --------------------------------------------------------------------------------
class TableEmbeddings:
    """Deterministic stand-in for the embedding model: text -> fixed random vector (the reference's
    tests use a MagicMock returning [0.1,0.2,0.3]*128, :44-50)."""

    def __init__(self, dim=384):
        self.dim, self.calls = dim, {"docs": 0, "query": 0}

    def _vec(self, text):
        import zlib
        rng = np.random.default_rng(zlib.crc32(text.encode()))
        v = rng.standard_normal(self.dim)
        return (v / np.linalg.norm(v)).astype(np.float32)

    def embed_documents(self, texts):
        self.calls["docs"] += 1
        return [self._vec(t).tolist() for t in texts]

    def embed_query(self, text):
        self.calls["query"] += 1
        return self._vec(text).tolist()


@pytest.fixture(params=[1, 2, 8], ids=["1gpu", "2gpus", "8gpus"])
def fresh_store(request):
    """The same suite runs on a single-GPU collection and on collections row-sharded over 2 and 8 GPUs of the
    box (``devices=``): the reference-facing behaviour must not depend on where the rows live."""
    import torch
    from archi_b200 import B200VectorStore
    n_dev = request.param
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    devices = list(range(n_dev)) if n_dev > 1 else None
    names = []

    def make(name, **kw):
        B200VectorStore.drop_collection(name, devices=devices)
        names.append(name)
        return B200VectorStore(pg_config={}, embedding_function=kw.pop("emb", TableEmbeddings()),
                               collection_name=name, devices=devices, **kw)
    make.devices = devices
    yield make
    for n in names:
        B200VectorStore.drop_collection(n, devices=devices)


def test_add_and_search_conventions(fresh_store):
    from archi_b200 import Document
    emb = TableEmbeddings()
    vs = fresh_store("t_conv", emb=emb)
    texts = [f"chunk number {i}" for i in range(50)]
    metas = [{"filename": f"f{i % 5}.md", "idx": i} for i in range(50)]
    ids = vs.add_texts(texts, metadatas=metas)
    assert len(ids) == 50 and len(set(ids)) == 50 and emb.calls["docs"] == 1     # one embed_documents call
    assert vs.count() == 50
    res = vs.similarity_search_with_score("chunk number 17", k=5)
    assert emb.calls["query"] == 1
    assert len(res) == 5
    doc, score = res[0]
    assert isinstance(doc, Document) and doc.page_content == "chunk number 17"
    assert score == pytest.approx(1.0, abs=1e-5)                                   # score = 1 - distance
    assert doc.metadata["collection"] == "t_conv" and doc.metadata["chunk_id"] == ids[17]
    assert doc.metadata["filename"] == "f2.md"
    assert [s for _, s in res] == sorted((s for _, s in res), reverse=True)
    # oracle on the same embeddings
    corpus = np.asarray(emb.embed_documents(texts), dtype=np.float32)
    q = np.asarray(emb.embed_query("chunk number 17"), dtype=np.float32)
    d, i = orc.exact_topk("cosine", corpus, q, 5)
    assert [r[0].metadata["idx"] for r in res] == i[0].tolist()
    assert np.allclose([r[1] for r in res], 1.0 - d[0], rtol=1e-5, atol=1e-6)
    assert [d.page_content for d in vs.similarity_search("chunk number 17", k=5)] == [r[0].page_content for r in res]
    assert vs.similarity_search_by_vector(q.tolist(), k=2)[0].page_content == "chunk number 17"
    # edge cases of :517-558 -- empty query, huge k, odd strings: a list comes back
    assert len(vs.similarity_search("", k=3)) == 3
    assert len(vs.similarity_search("chunk", k=10000)) == 50
    assert isinstance(vs.similarity_search("'; DROP TABLE document_chunks; --", k=2), list)
    assert isinstance(vs.similarity_search("日本語のクエリ 🚀", k=2), list)


@pytest.mark.parametrize("metric", ["l2", "inner_product"])
def test_distance_metrics_return_raw_distance(fresh_store, metric):
    emb = TableEmbeddings(64)
    vs = fresh_store(f"t_{metric}", emb=emb, distance_metric=metric)
    texts = [f"t{i}" for i in range(40)]
    vs.add_texts(texts)
    res = vs.similarity_search_with_score("t3", k=4)
    corpus = np.asarray(emb.embed_documents(texts), dtype=np.float32)
    d, i = orc.exact_topk(metric, corpus, np.asarray(emb.embed_query("t3"), dtype=np.float32), 4)
    assert [r[0].page_content for r in res] == [texts[j] for j in i[0]]
    assert np.allclose([r[1] for r in res], d[0], rtol=1e-5, atol=2e-6)            # ascending distance
    if metric == "inner_product":
        assert res[0][1] == pytest.approx(-1.0, abs=1e-5)                           # NEGATIVE inner product


def test_filter_delete_upsert_and_document_metadata(fresh_store):
    vs = fresh_store("t_filter")
    vs.register_document(7, resource_hash="abc123", display_name="Seven", source_type="web", url="http://x/7")
    vs.register_document(8, resource_hash="def456")
    vs.add_texts([f"seven {i}" for i in range(6)], metadatas=[{"kind": "a", "n": i} for i in range(6)], document_id=7)
    ids8 = vs.add_texts([f"eight {i}" for i in range(4)], metadatas=[{"kind": "b"} for _ in range(4)], document_id=8)
    vs.add_texts(["loose chunk"], metadatas=[None and {} or {}])
    assert vs.count() == 11
    res = vs.similarity_search_with_score("seven 2", k=20, filter={"kind": "b"})
    assert len(res) == 4 and all(r[0].metadata["kind"] == "b" and r[0].metadata["resource_hash"] == "def456" for r in res)
    res = vs.similarity_search("seven 2", k=3, filter={"kind": "a", "n": 2})      # values compared as text
    assert [d.page_content for d in res] == ["seven 2"]
    assert res[0].metadata["display_name"] == "Seven" and res[0].metadata["url"] == "http://x/7"
    assert vs.similarity_search("x", k=3, filter={"missing": "1"}) == []
    # soft-deleted documents are hidden unless include_deleted (:304-308)
    vs.register_document(8, is_deleted=True)
    assert all(d.metadata.get("kind") != "b" for d in vs.similarity_search("eight 1", k=11))
    assert vs.similarity_search("eight 1", k=1, include_deleted=True)[0].page_content == "eight 1"
    vs.register_document(8, is_deleted=False)
    # delete by chunk id and by document id (:493-535)
    assert vs.delete(ids=[ids8[0]]) is True and vs.count() == 10
    assert vs.similarity_search("eight 0", k=1)[0].page_content != "eight 0"
    assert vs.delete(document_id=7) is True and vs.count() == 4
    assert vs.delete() is False
    # upsert on (document_id, chunk_index) (:173-176)
    vs.add_texts(["eight one, revised"], document_id=8)                            # replaces chunk_index 0? no: index 0 was deleted
    vs.add_texts(["eight 1 v2", "eight 2 v2"], document_id=8)                      # indices 0,1 again -> replaces
    texts = sorted(d.page_content for d in vs.similarity_search("eight", k=50))
    assert "eight one, revised" not in texts and "eight 1 v2" in texts and "eight 1" not in texts
    # metadata is never None (:560-581)
    assert all(isinstance(d.metadata, dict) for d in vs.similarity_search("loose", k=50))


def test_store_objects_share_the_collection(fresh_store):
    from archi_b200 import B200VectorStore
    emb = TableEmbeddings(32)
    a = fresh_store("t_shared", emb=emb)
    a.add_texts(["alpha", "beta"])
    b = B200VectorStore(pg_config={}, embedding_function=emb, collection_name="t_shared",
                        devices=fresh_store.devices)                                          # per-request construction
    assert b.count() == 2 and b.similarity_search("beta", k=1)[0].page_content == "beta"
    with pytest.raises(ValueError, match="metric is fixed"):
        B200VectorStore(None, emb, "t_shared", "l2", devices=fresh_store.devices)


def test_hybrid_search_matches_oracle(fresh_store):
    from archi_b200 import HybridRetriever
    emb = TableEmbeddings(128)
    vs = fresh_store("t_hybrid", emb=emb)
    texts = ["the detector measures muon momentum", "jets are clustered with anti-kt", "muon chambers and muon triggers",
             "the trigger menu selects events", "grid computing sites run jobs", "muon", "tracker alignment constants"] * 3
    texts = [f"{t} #{i}" for i, t in enumerate(texts)]
    vs.add_texts(texts, metadatas=[{"i": i} for i in range(len(texts))])
    query = "muon trigger"
    corpus = np.asarray(emb.embed_documents(texts), dtype=np.float32)
    q = np.asarray(emb.embed_query(query), dtype=np.float32)
    bm = orc.bm25_scores([orc.tokenize(t) for t in texts], orc.tokenize(query))
    for ws, wb in ((0.7, 0.3), (0.4, 0.6), (0.0, 1.0)):
        res = vs.hybrid_search(query, k=6, semantic_weight=ws, bm25_weight=wb)
        comb, ids = orc.exact_hybrid_topk("cosine", corpus, q, bm, ws, wb, 6)
        assert [r[0].metadata["i"] for r in res] == ids.tolist()
        assert np.allclose([r[1] for r in res], comb, rtol=1e-5, atol=2e-6)
    r = HybridRetriever(vs, k=4, bm25_weight=0.6, semantic_weight=0.4)
    got = r.invoke(query)
    comb, ids = orc.exact_hybrid_topk("cosine", corpus, q, bm, 0.4, 0.6, 4)
    assert [g[0].metadata["i"] for g in got] == ids.tolist()


def test_hybrid_without_bm25_index_raises_and_retriever_reraises(fresh_store):
    from archi_b200 import HybridRetriever
    vs = fresh_store("t_nobm25", bm25_index=False)
    vs.add_texts(["a b c", "d e f"])
    with pytest.raises(RuntimeError, match="BM25 index"):           # :281-290
        vs.hybrid_search("a", k=1)
    with pytest.raises(RuntimeError, match="BM25 index"):           # hybrid_retriever.py:93-99
        HybridRetriever(vs).invoke("a")


def test_b200_embeddings_end_to_end(fresh_store):
    import torch
    from archi_b200 import B200Embeddings
    emb = B200Embeddings(dtype="f32")
    texts = ["The detector measures muon momentum.\nSecond line.", "Jets are clustered with anti-kt.", "short", "x " * 400]
    vecs = np.asarray(emb.embed_documents(texts), dtype=np.float32)
    assert vecs.shape == (4, 384)                                    # test_ingestion_pipeline_isolation.py:128-142
    assert np.allclose(np.linalg.norm(vecs, axis=1), 1.0, atol=1e-5)
    # the fused kernel equals torch's own mean-pool + normalise on the same hidden states
    hidden, mask = emb._forward(emb._clean(texts))
    m = mask.unsqueeze(-1).float()
    ref = (hidden.float() * m).sum(1) / m.sum(1).clamp(min=1e-9)
    ref = torch.nn.functional.normalize(ref, p=2, dim=1)
    assert np.allclose(vecs, ref.cpu().numpy(), rtol=1e-4, atol=1e-5)
    vs = fresh_store("t_e2e", emb=emb)
    vs.add_texts(texts)                                              # goes through pool_normalize_append
    assert vs.count() == 4
    stored = np.concatenate([sh.native.read_rows(0, len(sh.l2g)) for sh in vs._coll.shards if sh.l2g])
    order = np.concatenate([sh.l2g for sh in vs._coll.shards if sh.l2g])
    assert np.allclose(stored[np.argsort(order)], vecs, rtol=1e-5, atol=2e-6)
    top = vs.similarity_search_with_score(texts[1], k=1)[0]
    assert top[0].page_content == texts[1] and top[1] == pytest.approx(1.0, abs=1e-4)


def test_collection_snapshot_and_pgvector_import(fresh_store, tmp_path):
    """Restart story: save -> drop -> load gives the same answers (dense, filtered, hybrid, soft-deleted
    documents, tombstones), and a collection rebuilt from the reference's wire format (`embedding::text`) too."""
    from archi_b200 import B200VectorStore
    emb = TableEmbeddings(64)
    vs = fresh_store("t_snap", emb=emb)
    texts = [f"{w} chunk {i} about {'muon' if i % 3 == 0 else 'jets'} triggers" for i, w in enumerate(["alpha", "beta", "gamma", "delta"] * 10)]
    metas = [{"filename": f"f{i % 4}.md", "i": i} for i in range(40)]
    vs.add_texts(texts[:20], metas[:20], ids=[f"c{i}" for i in range(20)], document_id=1)
    vs.add_texts(texts[20:], metas[20:], ids=[f"c{i}" for i in range(20, 40)], document_id=2)
    vs.register_document(1, resource_hash="h1", display_name="One", source_type="local", url="u1")
    vs.register_document(2, resource_hash="h2", display_name="Two", source_type="web", url=None, is_deleted=True)
    vs.delete(ids=["c3", "c4"])

    def answers(store):
        out = [store.similarity_search_with_score("alpha chunk 8 about jets triggers", k=6),
               store.similarity_search_with_score("beta chunk 9", k=5, filter={"filename": "f1.md"}),
               store.similarity_search_with_score("gamma", k=50, include_deleted=True),
               store.hybrid_search("muon triggers", k=5, semantic_weight=0.4, bm25_weight=0.6)]
        return [[(d.page_content, sorted(d.metadata.items(), key=str), round(float(s), 6)) for d, s in r] for r in out], store.count()

    want = answers(vs)
    snap = str(tmp_path / "snap")
    vs.save(snap)
    devices = fresh_store.devices
    B200VectorStore.drop_collection("t_snap", devices=devices)
    back = B200VectorStore.load(snap, emb, pg_config={}, devices=devices)
    assert answers(back) == want
    back.add_texts(["a brand new chunk"], [{"filename": "new.md"}], ids=["n0"])            # the restored store keeps working
    assert back.similarity_search("a brand new chunk", k=1)[0].page_content == "a brand new chunk"
    B200VectorStore.drop_collection("t_snap", devices=devices)

    # the reference's table as it would be read back: embedding::text, metadata jsonb
    import json
    rows = []
    for i, t in enumerate(texts):
        if i in (3, 4):
            continue                                                         # DELETEd rows are not in the table
        v = emb._vec(t)
        md = dict(metas[i], collection="t_import", chunk_id=f"c{i}")
        rows.append((1 if i < 20 else 2, i % 20, t, "[" + ",".join(str(x) for x in v.tolist()) + "]", json.dumps(md)))
    rows.append((9, 0, "other collection row", "[" + ",".join(["0.5"] * 64) + "]", {"collection": "somewhere_else"}))
    imp = fresh_store("t_import", emb=emb)
    assert imp.import_pgvector_rows(rows) == 38
    imp.register_document(1, resource_hash="h1", display_name="One", source_type="local", url="u1")
    imp.register_document(2, resource_hash="h2", display_name="Two", source_type="web", url=None, is_deleted=True)
    got = answers(imp)
    strip = lambda res: [[(t, [kv for kv in md if kv[0] != "collection"], s) for t, md, s in r] for r in res]  # noqa: E731
    assert strip(got[0]) == strip(want[0]) and got[1] == want[1]
    assert imp.delete(ids=["c7"]) and imp.count() == 37
