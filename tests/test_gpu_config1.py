"""BASELINE config 1 end to end on the GPU: archi's docs (pseudonymised fixture, identical chunking) ->
split_text -> B200Embeddings (MiniLM-shaped encoder forward in PyTorch + fused pool/normalise/append kernel)
-> B200VectorStore.add_texts -> similarity_search(k=5), and the same files through IngestionDriver.add_files.

Checked against: the oracle's exact search over the rows the store holds (ids identical up to ties, scores
1e-5), the oracle's pool+normalise on the encoder's hidden states, and a full CPU restatement of the
reference path (torch CPU fp32 encoder -> oracle pool+normalise -> oracle exact search).
Reference: manager.py:75-78,292-324,362-373; postgres_vectorstore.py:143,207-248."""
import os
import time

import numpy as np
import pytest

from oracle import oracle as orc

import config1_data

pytestmark = pytest.mark.gpu


def _chunks_and_meta():
    from archi_b200.ingest import split_text
    fx = config1_data.load()
    chunks, metas = [], []
    for doc in fx["docs"]:
        for i, c in enumerate(split_text(doc["text"], 1000, 0, "\n\n")):
            chunks.append(c)
            metas.append({"filename": doc["filename"], "chunk_index": i, "resource_hash": "h-" + doc["filename"]})
    return fx, chunks, metas


def _cpu_embed(ef, texts):
    """The reference's embed path restated on the CPU: same weights, torch fp32, oracle pool + normalise."""
    import copy
    import torch
    model = copy.deepcopy(ef.model).to("cpu", torch.float32).eval()
    out = []
    texts = [t.replace("\n", " ") for t in texts]
    for s in range(0, len(texts), 32):
        ids, mask = ef.tokenizer(texts[s:s + 32], ef.max_seq_length)
        with torch.inference_mode():
            hidden = model(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask)).last_hidden_state
        out.append(orc.pool_normalize(hidden.numpy(), mask))
    return np.concatenate(out).astype(np.float32)


def test_config1_add_texts_and_top5(tmp_path):
    import torch
    from archi_b200 import B200VectorStore
    from archi_b200.embeddings import B200Embeddings
    from archi_b200.ingest import IngestionDriver
    fx, chunks, metas = _chunks_and_meta()
    assert len(chunks) == fx["n_chunks"] == 120
    ef = B200Embeddings(dtype="f32", seed=0, batch_size=32)       # sentence-transformers' batching: 32 texts per forward
    name = "config1_docs"
    B200VectorStore.drop_collection(name)
    store = B200VectorStore({}, ef, collection_name=name)
    t0 = time.perf_counter()
    ids = store.add_texts(chunks, [dict(m) for m in metas])
    torch.cuda.synchronize()
    t_add = time.perf_counter() - t0
    assert len(ids) == 120 and store.count() == 120
    stored = store.native.read_rows(0, 120)
    assert np.allclose(np.linalg.norm(stored, axis=1), 1.0, atol=1e-5)

    # the fused pool+normalise+append kernel against the oracle, on the encoder's own hidden states
    hidden, mask = ef._forward([c.replace("\n", " ") for c in chunks[:32]])
    want = orc.pool_normalize(hidden.float().cpu().numpy(), mask.cpu().numpy())
    assert np.allclose(stored[:32], want, rtol=1e-5, atol=2e-6)

    # the default token-budget batching (one forward for the 120 chunks, lengths padded to a multiple of 32) gives
    # the same rows up to the encoder's fp32 rounding across batch shapes
    ef_budget = B200Embeddings(dtype="f32", seed=0)
    assert ef_budget.batch_size is None
    rows_budget = ef_budget.embed_documents_device(chunks).cpu().numpy()
    assert rows_budget.shape == stored.shape and np.allclose(rows_budget, stored, atol=2e-5)
    del ef_budget

    # top-5 through the store surface against the oracle's exact search over the stored rows
    queries = config1_data.queries(chunks, 20)
    t0 = time.perf_counter()
    results = [store.similarity_search_with_score(q, k=5) for q in queries]
    t_search = time.perf_counter() - t0
    q_emb = np.asarray([ef.embed_query(q) for q in queries], dtype=np.float32)
    d_true, i_true = orc.exact_topk("cosine", stored, q_emb, 5)
    for qi, res in enumerate(results):
        assert len(res) == 5
        got_ids = [chunks.index(doc.page_content) for doc, _ in res]
        got_sc = np.asarray([[s for _, s in res]], dtype=np.float32)
        fails = orc.verify_topk("cosine", stored, q_emb[qi:qi + 1], 5, np.asarray([got_ids]), got_sc, 1e-5,
                                d_true[qi:qi + 1], i_true[qi:qi + 1])
        assert not fails, fails
        doc0 = res[0][0]
        assert doc0.metadata["filename"] == metas[got_ids[0]]["filename"] and doc0.metadata["collection"] == name
    # (how often a query quoting a chunk finds that chunk first depends on the encoder's weights, which are random
    #  here: reported, not asserted)
    step = max(1, len(chunks) // 20)
    quoted_first = sum(chunks.index(r[0][0].page_content) == i * step for i, r in enumerate(results))

    # the whole path restated on the CPU (encoder forward in torch fp32): same neighbours
    cpu_rows = _cpu_embed(ef, chunks)
    assert np.allclose(cpu_rows, stored, atol=2e-4)
    cpu_q = _cpu_embed(ef, queries)
    _, i_cpu = orc.exact_topk("cosine", cpu_rows, cpu_q, 5)
    overlap = np.mean([len(set(i_cpu[q]) & set(i_true[q])) / 5.0 for q in range(len(queries))])
    assert overlap >= 0.97, overlap

    # the same files through the ingestion driver (one length-ordered embedding pass for the 15 files)
    for doc in fx["docs"]:
        (tmp_path / doc["filename"]).write_text(doc["text"], encoding="utf-8")
    B200VectorStore.drop_collection(name + "_driver")
    store2 = B200VectorStore({}, ef, collection_name=name + "_driver")
    t0 = time.perf_counter()
    report = IngestionDriver(store2).add_files({"h-" + d["filename"]: str(tmp_path / d["filename"]) for d in fx["docs"]})
    torch.cuda.synchronize()
    t_ingest = time.perf_counter() - t0
    assert report.chunks == 120 and len(report.embedded) == 15 and not report.failed and report.embed_calls == 1
    stored2 = store2.native.read_rows(0, 120)
    assert np.allclose(stored2, stored, atol=1e-5)          # batch composition differs (length-ordered), values do not
    for q, res in zip(queries[:5], results[:5]):
        res2 = store2.similarity_search_with_score(q, k=5)
        assert [d.page_content for d, _ in res2] == [d.page_content for d, _ in res]
        assert res2[0][0].metadata["resource_hash"] == "h-" + res2[0][0].metadata["filename"]
    print(f"config1: add_texts {120 / t_add:.0f} chunks/s, driver {120 / t_ingest:.0f} chunks/s, "
          f"similarity_search {len(queries) / t_search:.0f} q/s (one query per call, k=5); "
          f"{quoted_first}/20 quoting queries found their chunk first (random-init encoder)")
    B200VectorStore.drop_collection(name)
    B200VectorStore.drop_collection(name + "_driver")
