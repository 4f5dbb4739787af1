"""Reference-driven conformance harness (TEST INFRASTRUCTURE).

Runs the UNMODIFIED reference classes -- /root/reference/src/data_manager/vectorstore/postgres_vectorstore.py
and retrievers/*.py, loaded by file path -- against a stand-in for the engines they talk to, and runs the same
scripted scenario against archi_b200.B200VectorStore, so that the two transcripts can be compared.

What is real and what is a stand-in:
  * real: every line of the reference's Python on this path (SQL construction, WHERE parameters, query-vector
    text serialisation :313/:391, score conversion :361, metadata merge :342-354, hybrid fallback :466-469,
    BM25-index RuntimeError :415-418, upsert statement :168-180, delete :493-535, count :570-585, retriever
    policies).  The module stubbing follows the reference's own unit test
    (tests/unit/test_vectorstore_manager_batch_commit.py:8-73).
  * stand-in: psycopg2 + PostgreSQL + pgvector + pg_textsearch.  ``FakePg`` executes exactly the statement
    shapes the reference emits (it parses the operator, the WHERE clauses and the parameter list out of the SQL
    text it receives) and computes distances / BM25 with the oracle (oracle.c float accumulators = restated
    pgvector; oracle.bm25_scores = restated pg_textsearch).  Those two engines remain "parity unpinned".

The reference tree does not travel to the GPU box: tests/golden/make_reference_golden.py records the
reference's transcript into tests/golden/reference_conformance.json in this container; the GPU test replays
the scenario on the real CUDA store and compares with that fixture.
"""
from __future__ import annotations

import copy
import importlib.util
import json
import os
import re
import sys
import types
import zlib
from typing import Any, Callable, Dict, List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_ROOT = "/root/reference"
REF_VS = os.path.join(REF_ROOT, "src/data_manager/vectorstore")
DIM = 48


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF_VS, "postgres_vectorstore.py"))


# ---------------------------------------------------------------------------------------------
# deterministic embedding function (LangChain Embeddings surface: lists of Python floats)
# ---------------------------------------------------------------------------------------------
class HashEmbeddings:
    """Bag-of-words embedding: sum of a fixed pseudo-random unit vector per token, normalised.  Texts that
    share words are close, so nearest-neighbour order is meaningful and well separated."""

    def __init__(self, dim: int = DIM):
        self.dim = dim
        self.calls: List[tuple] = []

    def _vec(self, text: str) -> np.ndarray:
        toks = re.findall(r"[a-z0-9]+", text.lower()) or ["<empty>"]
        v = np.zeros(self.dim, dtype=np.float64)
        for t in toks:
            v += np.random.default_rng(zlib.crc32(t.encode())).standard_normal(self.dim)
        v += 0.05 * np.random.default_rng(zlib.crc32(text.encode("utf-8")) ^ 0x5bd1e995).standard_normal(self.dim)
        return (v / np.linalg.norm(v)).astype(np.float32)

    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        self.calls.append(("embed_documents", len(texts)))
        return [self._vec(t).tolist() for t in texts]

    def embed_query(self, text: str) -> List[float]:
        self.calls.append(("embed_query", text))
        return self._vec(text).tolist()


class ScaledEmbeddings(HashEmbeddings):
    """Rows of different lengths, for the l2 / inner_product stores."""

    def _vec(self, text: str) -> np.ndarray:
        v = super()._vec(text)
        return (v * (0.5 + (zlib.crc32(text.encode("utf-8")) % 1000) / 1000.0)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# stand-in for psycopg2 + PostgreSQL/pgvector/pg_textsearch
# ---------------------------------------------------------------------------------------------
class RealDictCursor:      # sentinel, like psycopg2.extras.RealDictCursor
    pass


def _json_text(v: Any) -> Optional[str]:
    """metadata->>'key' for a JSON value."""
    if v is None:
        return None
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, str):
        return v
    return json.dumps(v)


class FakePg:
    """One database: document_chunks + documents (init.sql:230-308), a bm25 index flag."""

    def __init__(self, has_bm25_index: bool = True):
        from oracle import oracle as orc
        self.orc = orc
        self.chunks: List[Dict[str, Any]] = []
        self.documents: Dict[Any, Dict[str, Any]] = {}
        self.next_id = 1
        self.has_bm25_index = has_bm25_index
        self.statements: List[str] = []
        self.connections = 0

    # psycopg2.connect(**pg_config)
    def connect(self, **_cfg):
        self.connections += 1
        return _FakeConnection(self)

    # ---- engine pieces ---------------------------------------------------------------------------------
    def _distance(self, op: str, emb: np.ndarray, q: np.ndarray) -> float:
        metric = {"<=>": 0, "<->": 1, "<#>": 2}[op]
        lib = self.orc.clib()
        a = np.ascontiguousarray(emb, dtype=np.float32)
        b = np.ascontiguousarray(q, dtype=np.float32)
        return float(lib.orc_distance_f32(metric, a.size, a.ctypes.data, b.ctypes.data))

    def _where(self, sql: str, params: List[Any]):
        """Evaluate the reference's WHERE clause (:296-310): returns (rows passing, params consumed)."""
        keys = re.findall(r"c\.metadata->>'([^']+)' = %s", sql)
        assert keys and keys[0] == "collection" and "c.metadata->>'collection' IS NULL" in sql, sql
        vals = params[:len(keys)]
        check_deleted = "d.is_deleted = FALSE" in sql
        out = []
        for row in self.chunks:
            md = row["metadata"] or {}
            coll = _json_text(md.get("collection"))
            if not (coll == vals[0] or coll is None):
                continue
            ok = True
            for key, val in zip(keys[1:], vals[1:]):
                if _json_text(md.get(key)) != val:
                    ok = False
            if not ok:
                continue
            doc = self.documents.get(row["document_id"]) if row["document_id"] is not None else None
            if check_deleted and doc is not None and doc.get("is_deleted"):
                continue
            out.append((row, doc))
        return out, len(keys)

    @staticmethod
    def _doc_cols(doc):
        return {f: (doc.get(f) if doc else None) for f in ("resource_hash", "display_name", "source_type", "url")}


class _FakeConnection:
    encoding = "UTF8"

    def __init__(self, db: FakePg):
        self.db, self.closed, self.commits = db, False, 0

    def cursor(self, cursor_factory=None):
        return _FakeCursor(self.db, self, cursor_factory is RealDictCursor)

    def commit(self):
        self.commits += 1

    def close(self):
        self.closed = True


class _FakeCursor:
    def __init__(self, db: FakePg, conn: _FakeConnection, dict_rows: bool):
        self.db, self.connection, self.dict_rows = db, conn, dict_rows
        self._rows: List[Any] = []
        self.rowcount = -1

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def fetchall(self):
        rows, self._rows = self._rows, []
        return rows

    def fetchone(self):
        return self._rows.pop(0) if self._rows else None

    # psycopg2.extras.execute_values(cursor, sql, argslist, template=...)
    def execute_values(self, sql: str, argslist, template=None):
        flat = " ".join(sql.split())
        self.db.statements.append(flat)
        assert flat.startswith("INSERT INTO document_chunks (document_id, chunk_index, chunk_text, embedding, metadata)"), flat
        assert "ON CONFLICT (document_id, chunk_index) DO UPDATE SET" in flat and template == "(%s, %s, %s, %s::vector, %s::jsonb)"
        for document_id, chunk_index, text, embedding, metadata_json in argslist:
            emb = np.asarray(embedding, dtype=np.float32)          # ::vector stores float4
            md = json.loads(metadata_json)
            hit = None
            if document_id is not None:                             # NULL never conflicts in a UNIQUE constraint
                hit = next((r for r in self.db.chunks if r["document_id"] == document_id and r["chunk_index"] == chunk_index), None)
            if hit is not None:
                hit.update(chunk_text=text, embedding=emb, metadata=md)
            else:
                self.db.chunks.append(dict(id=self.db.next_id, document_id=document_id, chunk_index=chunk_index,
                                           chunk_text=text, embedding=emb, metadata=md))
                self.db.next_id += 1

    def execute(self, sql: str, params=None):
        flat = " ".join(sql.split())
        self.db.statements.append(flat)
        params = list(params) if params is not None else []
        db = self.db
        if "FROM pg_class t" in flat and "am.amname = 'bm25'" in flat:
            self._rows = [{"relname": "idx_document_chunks_bm25"}] if db.has_bm25_index else []
        elif flat.startswith("WITH scored AS"):
            self._hybrid(flat, params)
        elif " AS distance" in flat:
            self._semantic(flat, params)
        elif flat.startswith("DELETE FROM document_chunks WHERE document_id = %s"):
            before = len(db.chunks)
            db.chunks = [r for r in db.chunks if r["document_id"] != params[0]]
            self.rowcount = before - len(db.chunks)
        elif flat.startswith("DELETE FROM document_chunks WHERE metadata->>'chunk_id' = %s"):
            before = len(db.chunks)
            db.chunks = [r for r in db.chunks if _json_text((r["metadata"] or {}).get("chunk_id")) != params[0]]
            self.rowcount = before - len(db.chunks)
        elif flat.startswith("SELECT COUNT(*) FROM document_chunks"):
            n = sum(1 for r in db.chunks if _json_text((r["metadata"] or {}).get("collection")) in (params[0], None))
            self._rows = [(n,)]
        else:
            raise AssertionError("FakePg: unexpected statement: " + flat)

    def _parse_vector(self, text: str) -> np.ndarray:
        assert text.startswith("[") and text.endswith("]")
        return np.asarray([float(x) for x in text[1:-1].split(",")], dtype=np.float64).astype(np.float32)

    def _semantic(self, flat: str, params: List[Any]):
        op = re.search(r"c\.embedding (<=>|<->|<#>) %s::vector AS distance", flat).group(1)
        assert "ORDER BY distance ASC LIMIT %s" in flat and "LEFT JOIN documents d ON c.document_id = d.id" in flat
        q = self._parse_vector(params[0])
        rows, used = self.db._where(flat, params[1:])
        k = params[1 + used]
        assert len(params) == 2 + used
        scored = []
        for row, doc in rows:
            d = self.db._distance(op, row["embedding"], q)
            scored.append((d, row["id"], row, doc))
        scored.sort(key=lambda t: (np.inf if np.isnan(t[0]) else t[0], t[1]))
        self._rows = [dict(id=row["id"], chunk_text=row["chunk_text"], metadata=copy.deepcopy(row["metadata"]),
                           distance=d, **FakePg._doc_cols(doc)) for d, _, row, doc in scored[:k]]

    def _hybrid(self, flat: str, params: List[Any]):
        op = re.search(r"1\.0 - \(c\.embedding (<=>|<->|<#>) %s::vector\) AS semantic_score", flat).group(1)
        assert re.search(r"c\.chunk_text <@> to_bm25query\(%s, 'idx_document_chunks_bm25'\) AS bm25_score", flat), flat
        assert "(semantic_score * %s + COALESCE(bm25_score, 0) * %s) AS combined_score" in flat
        assert "ORDER BY combined_score DESC LIMIT %s" in flat
        q = self._parse_vector(params[0])
        rows, used = self.db._where(flat, params[1:])
        query_text, w_s, w_b, k = params[1 + used:]
        orc = self.db.orc
        # the BM25 index covers the whole table: N, df and avgdl come from every chunk
        docs_tokens = [orc.tokenize(r["chunk_text"]) for r in self.db.chunks]
        bm = orc.bm25_scores(docs_tokens, orc.tokenize(query_text))
        bm_by_id = {r["id"]: bm[i] for i, r in enumerate(self.db.chunks)}
        scored = []
        for row, doc in rows:
            sem = 1.0 - self.db._distance(op, row["embedding"], q)
            b = bm_by_id[row["id"]]
            b_sql = None if np.isnan(b) else float(b)
            combined = sem * w_s + (0.0 if b_sql is None else b_sql) * w_b
            scored.append((combined, row["id"], row, doc, sem, b_sql))
        scored.sort(key=lambda t: (-t[0], t[1]))
        self._rows = [dict(id=row["id"], chunk_text=row["chunk_text"], metadata=copy.deepcopy(row["metadata"]),
                           semantic_score=sem, bm25_score=b, combined_score=c, **FakePg._doc_cols(doc))
                      for c, _, row, doc, sem, b in scored[:k]]


# ---------------------------------------------------------------------------------------------
# loading the unmodified reference modules
# ---------------------------------------------------------------------------------------------
class _Document:
    def __init__(self, page_content: str, metadata: Optional[Dict[str, Any]] = None, **kw):
        self.page_content = page_content
        self.metadata = metadata if metadata is not None else {}


class _BaseRetriever:
    """Stand-in for langchain_core.retrievers.BaseRetriever (a pydantic model): keyword fields become
    attributes, ``invoke`` calls ``_get_relevant_documents``."""

    def __init__(self, **kwargs):
        for name, value in kwargs.items():
            setattr(self, name, value)

    def invoke(self, query, config=None, **kwargs):
        return self._get_relevant_documents(query)


_CURRENT_DB: Dict[str, FakePg] = {}


def _install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    def connect(**cfg):
        return _CURRENT_DB[cfg.get("dbname", "default")].connect(**cfg)

    extras = mod("psycopg2.extras", RealDictCursor=RealDictCursor,
                 execute_values=lambda cursor, sql, argslist, template=None, **kw: cursor.execute_values(sql, argslist, template))
    extensions = mod("psycopg2.extensions", connection=_FakeConnection)
    mod("psycopg2", connect=connect, extras=extras, extensions=extensions)
    mod("langchain_core")
    mod("langchain_core.documents", Document=_Document)
    mod("langchain_core.embeddings", Embeddings=object)
    mod("langchain_core.vectorstores", VectorStore=object)
    mod("langchain_core.vectorstores.base", VectorStore=object)
    mod("langchain_core.retrievers", BaseRetriever=_BaseRetriever)
    mod("langchain_core.callbacks")
    mod("langchain_core.callbacks.manager", CallbackManagerForRetrieverRun=object)


def _load_by_path(dotted: str, path: str):
    spec = importlib.util.spec_from_file_location(dotted, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = module
    spec.loader.exec_module(module)
    return module


_REF = None


def load_reference():
    """The reference's own classes, executed from the files under /root/reference (never copied)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    _install_stubs()
    # package shells so that `from src.utils.logging import get_logger` resolves without running
    # src/utils/__init__.py (which imports the whole application)
    for pkg in ("src", "src.utils", "src.data_manager", "src.data_manager.vectorstore", "src.data_manager.vectorstore.retrievers"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    _load_by_path("src.utils.logging", os.path.join(REF_ROOT, "src/utils/logging.py"))
    pv = _load_by_path("src.data_manager.vectorstore.postgres_vectorstore", os.path.join(REF_VS, "postgres_vectorstore.py"))
    _load_by_path("src.data_manager.vectorstore.retrievers.utils", os.path.join(REF_VS, "retrievers/utils.py"))
    hy = _load_by_path("src.data_manager.vectorstore.retrievers.hybrid_retriever", os.path.join(REF_VS, "retrievers/hybrid_retriever.py"))
    se = _load_by_path("src.data_manager.vectorstore.retrievers.semantic_retriever", os.path.join(REF_VS, "retrievers/semantic_retriever.py"))
    gr = _load_by_path("src.data_manager.vectorstore.retrievers.grading_retriever", os.path.join(REF_VS, "retrievers/grading_retriever.py"))
    _REF = types.SimpleNamespace(PostgresVectorStore=pv.PostgresVectorStore, HybridRetriever=hy.HybridRetriever,
                                 SemanticRetriever=se.SemanticRetriever, GradingRetriever=gr.GradingRetriever,
                                 Document=_Document)
    return _REF


# ---------------------------------------------------------------------------------------------
# the two implementations behind one small driver interface
# ---------------------------------------------------------------------------------------------
class ReferenceImpl:
    name = "reference"

    def __init__(self):
        self.ref = load_reference()
        self.Document = self.ref.Document
        self.HybridRetriever, self.SemanticRetriever, self.GradingRetriever = (
            self.ref.HybridRetriever, self.ref.SemanticRetriever, self.ref.GradingRetriever)
        _CURRENT_DB.clear()

    def database(self, name: str, bm25_index: bool = True):
        _CURRENT_DB[name] = FakePg(bm25_index)

    def store(self, db: str, collection: str, metric: str, emb):
        return self.ref.PostgresVectorStore({"dbname": db}, emb, collection_name=collection, distance_metric=metric)

    def from_texts(self, db, collection, metric, emb, texts, metadatas, **kw):
        return self.ref.PostgresVectorStore.from_texts(texts, emb, metadatas=metadatas, pg_config={"dbname": db},
                                                       collection_name=collection, distance_metric=metric, **kw)

    def register_document(self, db: str, store, document_id, **fields):
        _CURRENT_DB[db].documents[document_id] = dict(fields)

    def close(self):
        _CURRENT_DB.clear()


class B200Impl:
    name = "archi_b200"

    def __init__(self, patch: Optional[Callable] = None):
        import archi_b200.retrievers as r
        import archi_b200.vectorstore as vs
        self.vs = vs
        self.Document = vs.Document
        self.HybridRetriever, self.SemanticRetriever, self.GradingRetriever = r.HybridRetriever, r.SemanticRetriever, r.GradingRetriever
        self._bm25: Dict[str, bool] = {}
        self._collections: List[Any] = []

    def database(self, name: str, bm25_index: bool = True):
        self._bm25[name] = bm25_index

    def store(self, db: str, collection: str, metric: str, emb):
        cfg = {"host": "conformance", "dbname": db}
        if (collection, cfg) not in self._collections:
            self._collections.append((collection, cfg))
        return self.vs.B200VectorStore(cfg, emb, collection_name=collection, distance_metric=metric,
                                       bm25_index=self._bm25[db])

    def from_texts(self, db, collection, metric, emb, texts, metadatas, **kw):
        s = self.store(db, collection, metric, emb)
        s.add_texts(texts, metadatas=metadatas, **kw)
        return s

    def register_document(self, db: str, store, document_id, **fields):
        store.register_document(document_id, **fields)

    def close(self):
        for name, cfg in self._collections:
            self.vs.B200VectorStore.drop_collection(name, pg_config=cfg)


# ---------------------------------------------------------------------------------------------
# the scenario
# ---------------------------------------------------------------------------------------------
WORDS = ("muon detector trigger calorimeter tracker luminosity pileup jet electron photon higgs boson quark gluon "
         "neutrino cross section decay vertex momentum energy beam collider magnet cryostat readout firmware "
         "dataset workflow grid tier transfer quota ticket shift operator alarm voltage cooling").split()


def _texts(n: int, seed: int) -> List[str]:
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        m = int(rng.integers(5, 14))
        out.append(" ".join(rng.choice(WORDS, m).tolist()) + f" note{seed}x{i}")
    return out


def _results(res, chunk_ids: Optional[Dict[str, int]] = None):
    """(Document, score) list or Document list -> JSON-able rows; uuid chunk ids are replaced by their ordinal."""
    rows = []
    for item in res:
        doc, score = (item if isinstance(item, tuple) else (item, None))
        md = dict(doc.metadata)
        if chunk_ids is not None and md.get("chunk_id") in chunk_ids:
            md["chunk_id"] = f"<generated {chunk_ids[md['chunk_id']]}>"
        rows.append({"text": doc.page_content, "metadata": md, "score": None if score is None else float(score)})
    return rows


def _try(fn):
    try:
        return fn()
    except Exception as e:      # the transcript records the exception type and message
        return {"raises": type(e).__name__, "message": str(e)}


def run_scenario(impl) -> Dict[str, Any]:
    """Every public method of the store surface + the three retrievers; returns the transcript."""
    T: Dict[str, Any] = {}
    emb = HashEmbeddings()
    impl.database("main")
    impl.database("nobm25", bm25_index=False)
    Document = impl.Document

    # ---- constructor ----------------------------------------------------------------------------------
    T["ctor_bad_metric"] = _try(lambda: impl.store("main", "physics", "manhattan", emb))
    st = impl.store("main", "physics", "cosine", emb)
    T["ctor_attrs"] = [st._collection_name, st._distance_metric, st._distance_op, st.embeddings is emb]
    T["count_empty"] = st.count()
    T["search_empty"] = _results(st.similarity_search_with_score("muon trigger", k=3))
    T["hybrid_empty"] = _results(_try(lambda: st.hybrid_search("muon trigger", k=3)))
    T["add_nothing"] = st.add_texts([])

    # ---- add -------------------------------------------------------------------------------------------
    t1, t2, t3 = _texts(20, 1), _texts(20, 2), _texts(10, 3)
    md1 = [{"filename": f"a{i}.md", "kind": "a" if i % 2 else "b", "page": i % 4, "flag": bool(i % 3 == 0)} for i in range(20)]
    md2 = [{"filename": f"b{i}.md", "kind": "b", "page": i % 4, "nested": {"x": i}} for i in range(20)]
    T["add1_ids"] = st.add_texts(t1, md1, ids=[f"c{i}" for i in range(20)], document_id=1)
    T["add1_mutates_metadata"] = md1[3]
    T["add2_ids"] = st.add_texts(t2, md2, ids=[f"d{i}" for i in range(20)], document_id=2)
    gen_ids = st.add_documents([Document(page_content=t, metadata={"filename": f"c{i}.md", "kind": "c"}) for i, t in enumerate(t3)])
    T["add_documents_n_ids"] = [len(gen_ids), len(set(gen_ids)), all(isinstance(x, str) and len(x) == 36 for x in gen_ids)]
    gen = {cid: i for i, cid in enumerate(gen_ids)}
    impl.register_document("main", st, 1, resource_hash="hash-1", display_name="Doc One", source_type="local_files", url="https://x/1")
    impl.register_document("main", st, 2, resource_hash="hash-2", display_name="Doc Two", source_type="web", url=None)
    T["count_50"] = st.count()
    T["embed_calls_after_add"] = [c[0] for c in emb.calls]

    # ---- semantic search ---------------------------------------------------------------------------------
    q1, q2 = "muon trigger luminosity", "grid transfer quota ticket"
    R = lambda res: _results(res, gen)
    T["sim_with_score"] = R(st.similarity_search_with_score(q1, k=5))
    T["sim_docs_only"] = R(st.similarity_search(q2, k=3))
    T["sim_default_k"] = R(st.similarity_search_with_score(q2))
    vec = emb.embed_query("calorimeter readout firmware")
    T["by_vector"] = R(st.similarity_search_by_vector(vec, k=4))
    T["by_vector_with_score"] = R(st.similarity_search_by_vector_with_score(vec, k=4))
    T["k_10000"] = R(st.similarity_search_with_score(q1, k=10000))
    T["empty_query"] = R(st.similarity_search_with_score("", k=2))
    T["unicode_query"] = R(st.similarity_search_with_score("müon détecteur 粒子", k=2))
    T["injection_query"] = R(st.similarity_search_with_score("'; DROP TABLE document_chunks; --", k=2))
    # ---- filters (:296-310) ------------------------------------------------------------------------------
    T["filter_kind_a"] = R(st.similarity_search_with_score(q1, k=6, filter={"kind": "a"}))
    T["filter_page_int"] = R(st.similarity_search_with_score(q1, k=6, filter={"page": 3}))
    T["filter_two_keys"] = R(st.similarity_search_with_score(q1, k=6, filter={"kind": "b", "page": 1}))
    T["filter_bool_python_str"] = R(st.similarity_search_with_score(q1, k=6, filter={"flag": True}))
    T["filter_bool_json_str"] = R(st.similarity_search_with_score(q1, k=6, filter={"flag": "true"}))
    T["filter_missing_key"] = R(st.similarity_search_with_score(q1, k=6, filter={"nope": "x"}))
    T["filter_nested"] = R(st.similarity_search_with_score(q1, k=3, filter={"nested": '{"x": 4}'}))
    T["filter_empty_dict"] = R(st.similarity_search_with_score(q1, k=3, filter={}))
    # ---- soft-deleted documents (:304-308) ------------------------------------------------------------------
    impl.register_document("main", st, 2, resource_hash="hash-2", display_name="Doc Two", source_type="web", url=None, is_deleted=True)
    T["doc2_deleted"] = R(st.similarity_search_with_score(q2, k=8))
    T["doc2_include_deleted"] = R(st.similarity_search_with_score(q2, k=8, include_deleted=True))
    T["count_ignores_soft_delete"] = st.count()
    impl.register_document("main", st, 2, resource_hash="hash-2", display_name="Doc Two", source_type="web", url=None, is_deleted=False)

    # ---- hybrid (:366-491) -----------------------------------------------------------------------------------
    T["hybrid_default"] = R(st.hybrid_search(q1, k=5))
    T["hybrid_06_04"] = R(st.hybrid_search(q2, k=5, semantic_weight=0.4, bm25_weight=0.6))
    T["hybrid_filter"] = R(st.hybrid_search(q1, k=4, filter={"kind": "b"}))
    T["hybrid_no_lexical_match"] = R(st.hybrid_search("zzzz qqqq", k=3))
    T["hybrid_repeated_terms"] = R(st.hybrid_search("muon muon muon trigger", k=4))
    T["hybrid_k_10000_len"] = len(st.hybrid_search(q1, k=10000))
    T["hybrid_filter_no_rows_falls_back"] = R(st.hybrid_search(q1, k=3, filter={"nope": "x"}))
    nob = impl.store("nobm25", "physics", "cosine", emb)
    nob.add_texts(_texts(5, 9), ids=[f"n{i}" for i in range(5)])
    T["hybrid_without_index"] = _try(lambda: nob.hybrid_search(q1, k=3))

    # ---- a second store object on the same collection, another collection in the same database --------------
    st_again = impl.store("main", "physics", "cosine", emb)
    T["second_object_same_rows"] = [st_again.count(), R(st_again.similarity_search_with_score(q1, k=2))]
    other = impl.store("main", "other", "cosine", emb)
    other.add_texts(_texts(6, 4), [{"filename": f"o{i}.md"} for i in range(6)], ids=[f"o{i}" for i in range(6)])
    T["other_collection"] = [other.count(), R(other.similarity_search_with_score(q1, k=10))]
    T["physics_unaffected"] = st.count()

    # ---- upsert on (document_id, chunk_index) (:168-180) --------------------------------------------------------
    t1b = _texts(20, 5)
    st.add_texts(t1b, [{"filename": f"a{i}.v2.md", "kind": "a"} for i in range(20)], ids=[f"e{i}" for i in range(20)], document_id=1)
    T["upsert_count"] = st.count()
    T["upsert_search"] = R(st.similarity_search_with_score(t1b[7], k=3))
    T["upsert_old_text_gone"] = R(st.similarity_search_with_score(t1[7], k=1))

    # ---- delete (:493-535) ----------------------------------------------------------------------------------------
    T["delete_nothing"] = st.delete()
    T["delete_ids"] = st.delete(ids=["d0", "d1", "d2", "unknown"])
    T["count_after_delete_ids"] = st.count()
    T["delete_generated_id"] = st.delete(ids=[gen_ids[0]])
    T["delete_document"] = st.delete(document_id=2)
    T["count_after_delete_document"] = st.count()
    T["search_after_deletes"] = R(st.similarity_search_with_score(q2, k=6))
    T["hybrid_after_deletes"] = R(st.hybrid_search(q2, k=4, semantic_weight=0.4, bm25_weight=0.6))
    T["delete_empty_list"] = st.delete(ids=[])

    # ---- the other two metrics (:74-78, :361) ------------------------------------------------------------------------
    semb = ScaledEmbeddings()
    for metric in ("l2", "inner_product"):
        impl.database(metric)
        ms = impl.store(metric, "m", metric, semb)
        ms.add_texts(_texts(30, 6), [{"filename": f"m{i}.md"} for i in range(30)], ids=[f"m{i}" for i in range(30)])
        T[f"{metric}_attrs"] = [ms._distance_metric, ms._distance_op]
        T[f"{metric}_search"] = _results(ms.similarity_search_with_score(q1, k=5))
        T[f"{metric}_hybrid"] = _results(ms.hybrid_search(q1, k=5))

    # ---- from_texts -----------------------------------------------------------------------------------------------------
    impl.database("ft")
    ft = impl.from_texts("ft", "made", "cosine", emb, _texts(8, 7), [{"filename": f"f{i}.md"} for i in range(8)],
                         ids=[f"f{i}" for i in range(8)])
    T["from_texts"] = [ft.count(), _results(ft.similarity_search_with_score(q2, k=2))]

    # ---- retrievers ----------------------------------------------------------------------------------------------------------
    hr = impl.HybridRetriever(vectorstore=st, k=4, bm25_weight=0.6, semantic_weight=0.4)
    T["hybrid_retriever"] = R(hr.invoke(q1))
    hr_default = impl.HybridRetriever(vectorstore=st)
    T["hybrid_retriever_defaults"] = [hr_default.k, hr_default.bm25_weight, hr_default.semantic_weight, R(hr_default.invoke(q2))]
    T["hybrid_retriever_no_index_reraises"] = _try(lambda: impl.HybridRetriever(vectorstore=nob, k=2).invoke(q1))

    class _NoHybrid:           # a backend without hybrid_search: semantic-only fallback
        def __init__(self, inner):
            self.similarity_search_with_score = inner.similarity_search_with_score

    T["hybrid_retriever_semantic_fallback"] = R(impl.HybridRetriever(vectorstore=_NoHybrid(st), k=3).invoke(q1))

    class _Unsupported:
        def __init__(self, inner):
            self.similarity_search_with_score = inner.similarity_search_with_score

        def hybrid_search(self, **kw):
            raise RuntimeError("hybrid search is not supported by this backend")

    T["hybrid_retriever_unsupported_falls_back"] = R(impl.HybridRetriever(vectorstore=_Unsupported(st), k=3).invoke(q1))
    dm = {"embedding_name": "HuggingFaceEmbeddings",
          "embedding_class_map": {"HuggingFaceEmbeddings": {"kwargs": {"model_name": "sentence-transformers/all-MiniLM-L6-v2"}},
                                  "Qwen": {"kwargs": {"model": "Qwen/Qwen3-Embedding-0.6B"}}}}
    T["semantic_retriever"] = R(impl.SemanticRetriever(st, dm, k=3).invoke(q1))
    T["semantic_retriever_instructions_ignored"] = R(impl.SemanticRetriever(st, dm, k=2, instructions="find docs").invoke(q1))
    n_before = len(emb.calls)
    T["semantic_retriever_instructions_qwen"] = R(impl.SemanticRetriever(st, dict(dm, embedding_name="Qwen"), k=2,
                                                                         instructions="find docs").invoke(q1))
    T["instruction_query_text"] = emb.calls[n_before][1]
    T["grading_retriever"] = R(impl.GradingRetriever(st, k=3).invoke(q2))
    impl.close()
    return T


# ---------------------------------------------------------------------------------------------
# transcript comparison
# ---------------------------------------------------------------------------------------------
def compare(want: Any, got: Any, path: str = "", rel: float = 1e-5, fails: Optional[List[str]] = None) -> List[str]:
    """Structural equality; floats within ``rel`` (plus 1e-6 absolute).  Returns the list of differences."""
    fails = [] if fails is None else fails
    if isinstance(want, dict) and isinstance(got, dict):
        for k in sorted(set(want) | set(got)):
            if k not in want or k not in got:
                fails.append(f"{path}/{k}: present in only one transcript")
            else:
                compare(want[k], got[k], f"{path}/{k}", rel, fails)
    elif isinstance(want, (list, tuple)) and isinstance(got, (list, tuple)):
        if len(want) != len(got):
            fails.append(f"{path}: length {len(want)} vs {len(got)}")
        else:
            for i, (a, b) in enumerate(zip(want, got)):
                compare(a, b, f"{path}[{i}]", rel, fails)
    elif isinstance(want, float) and isinstance(got, (float, int)) and not isinstance(got, bool):
        if not abs(want - got) <= 1e-6 + rel * abs(want):
            fails.append(f"{path}: {want!r} vs {got!r}")
    elif want != got or type(want) is not type(got):
        fails.append(f"{path}: {want!r} vs {got!r}")
    return fails
