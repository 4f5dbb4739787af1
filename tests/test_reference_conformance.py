"""Conformance of B200VectorStore + retrievers with the reference's own classes.

tests/golden/reference_conformance.json is the transcript the UNMODIFIED reference PostgresVectorStore /
HybridRetriever / SemanticRetriever / GradingRetriever produced for tests/ref_harness.py::run_scenario
(recorded by tests/golden/make_reference_golden.py in the build container, where /root/reference exists).

  * test_fixture_is_what_the_reference_produces  (CPU, only where /root/reference exists): re-runs the
    reference and requires the committed fixture to be identical -- the fixture cannot drift.
  * test_host_logic_matches_reference            (CPU): the scenario on B200VectorStore with a numpy stand-in
    for the native store (TEST ONLY: the product has no CPU path) -- ids, metadata stamping, filters, upsert,
    delete, fallbacks, exceptions, retriever policies.
  * test_cuda_store_matches_reference            (-m gpu): the scenario on the real store through the C ABI.
Scores must agree within 1e-5 relative (fp32 storage); documents, metadata, ordering, lengths, return values
and exception types / messages exactly.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

from oracle import oracle as orc  # noqa: E402

FIXTURE = os.path.join(HERE, "golden", "reference_conformance.json")


def _fixture():
    return json.load(open(FIXTURE))["transcript"]


def _jsonable(t):
    return json.loads(json.dumps(t, ensure_ascii=False))


@pytest.mark.skipif(not H.reference_available(), reason="/root/reference is not on this machine")
def test_fixture_is_what_the_reference_produces():
    got = _jsonable(H.run_scenario(H.ReferenceImpl()))
    fails = H.compare(_fixture(), got, rel=0.0)
    assert not fails, fails[:10]


@pytest.mark.skipif(not H.reference_available(), reason="/root/reference is not on this machine")
def test_reference_emits_the_statements_the_stand_in_expects():
    """The stand-in database asserts on every statement shape; this pins the ones the design relies on:
    the query vector travels as decimal text (:313), a connection is opened per call (:94-98)."""
    ref = H.load_reference()
    H._CURRENT_DB.clear()
    H._CURRENT_DB["default"] = db = H.FakePg()
    emb = H.HashEmbeddings()
    st = ref.PostgresVectorStore({"dbname": "default"}, emb)
    st.add_texts(["muon trigger", "grid transfer"], ids=["a", "b"])
    st.similarity_search_with_score("muon", k=1)
    st.hybrid_search("muon", k=1)
    assert db.connections == 3
    kinds = [s.split(" ", 2)[0] for s in db.statements]
    assert kinds == ["INSERT", "SELECT", "SELECT", "WITH"]
    H._CURRENT_DB.clear()


# ---- CPU leg: numpy stand-ins for the native store and the BM25 kernel (tests only) ---------------------------
class _StandInNative:
    """What archi_b200.store.NativeStore does, restated with the oracle (fp64 truth)."""

    def __init__(self, dim, metric="cosine", storage_dtype="f32", device=0, capacity_rows=0):
        self.dim, self.metric, self.storage_dtype, self.device = dim, metric, storage_dtype, device
        self.data = np.empty((0, dim), dtype=np.float32)
        self.alive = np.empty(0, dtype=bool)

    def append(self, rows):
        a = np.asarray(rows, dtype=np.float32)
        first = self.data.shape[0]
        self.data = np.concatenate([self.data, a])
        self.alive = np.concatenate([self.alive, np.ones(a.shape[0], dtype=bool)])
        return first

    def delete_rows(self, rows):
        self.alive[np.asarray(rows, dtype=np.int64)] = False

    def count(self):
        return int(self.alive.sum())

    def rows(self):
        return self.data.shape[0]

    def close(self):
        pass

    def hybrid_search_terms(self, lexical, text_queries, queries, k, semantic_weight, bm25_weight, filter_mask=None,
                            include_deleted=False, id_offset=0):
        from test_lexical_host import bm25_from_csr
        bm = np.stack([np.nan_to_num(bm25_from_csr(lexical, tq), nan=0.0) for tq in text_queries])
        return self.search(queries, k, filter_mask=filter_mask, include_deleted=include_deleted, bm25=bm,
                           semantic_weight=semantic_weight, bm25_weight=bm25_weight, hybrid=True)

    def search(self, queries, k, filter_mask=None, include_deleted=False, bm25=None, semantic_weight=1.0,
               bm25_weight=0.0, hybrid=False, **_):
        q = np.atleast_2d(np.asarray(queries, dtype=np.float32))
        mask = self.alive.copy() if not include_deleted else np.ones_like(self.alive)
        if filter_mask is not None:
            bits = np.unpackbits(np.asarray(filter_mask).view(np.uint8), bitorder="little")[:mask.size].astype(bool)
            mask &= bits
        scores = np.full((q.shape[0], k), np.nan, dtype=np.float32)
        ids = np.full((q.shape[0], k), -1, dtype=np.int64)
        for i in range(q.shape[0]):
            if hybrid:
                b = None if bm25 is None else np.where(bm25[i] != 0, bm25[i], np.nan).astype(np.float64)
                s, r = orc.exact_hybrid_topk(self.metric, self.data, q[i], b, semantic_weight, bm25_weight, k, mask=mask)
            else:
                d, r = orc.exact_topk(self.metric, self.data, q[i], k, mask=mask)
                s, r = orc.score_from_distance(self.metric, d[0]), r[0]
            scores[i, :len(r)] = s
            ids[i, :len(r)] = r
        return scores, ids


@pytest.fixture
def cpu_stand_ins(monkeypatch):
    import archi_b200.vectorstore as vs
    from archi_b200.bm25 import LexicalIndex
    from test_lexical_host import bm25_from_csr
    monkeypatch.setattr(vs, "NativeStore", _StandInNative)
    monkeypatch.setattr(vs, "_upload_mask_words", lambda words, device: words)
    monkeypatch.setattr(LexicalIndex, "score",
                        lambda self, query, out=None: np.nan_to_num(bm25_from_csr(self, query), nan=0.0)[None, :])


def test_host_logic_matches_reference(cpu_stand_ins):
    got = _jsonable(H.run_scenario(H.B200Impl()))
    fails = H.compare(_fixture(), got, rel=1e-5)
    assert not fails, fails[:10]


@pytest.mark.gpu
def test_cuda_store_matches_reference():
    got = _jsonable(H.run_scenario(H.B200Impl()))
    fails = H.compare(_fixture(), got, rel=1e-5)
    assert not fails, fails[:10]
