"""B200VectorStore -- drop-in for archi's PostgresVectorStore
(reference: src/data_manager/vectorstore/postgres_vectorstore.py:25-585).

Same constructor, same method set, same score conventions and error behaviour; the embeddings
live in GPU HBM (archi_b200.store.NativeStore over libarchi_b200.so) instead of a pgvector column,
chunk text and metadata stay on the host.  There is no CPU search path: every search, hybrid
search and BM25 scoring call goes through the C ABI.

Differences that a maintainer should know (all deliberate):
  * search is always exact (the reference's default HNSW index may approximate the semantic path);
  * the reference builds a new store object per request (archi.py:61-65,
    vectorstore_connector.py:60-73); here objects with the same ``collection_name`` share one
    GPU-resident collection through a process-level registry, so construction stays cheap;
  * row ids are dense insertion indices, not a SERIAL column; ``chunk_id`` in metadata is kept.
"""
from __future__ import annotations

import json
import threading
import uuid
from typing import Any, Dict, Iterable, List, Optional, Tuple, Type

import numpy as np

from .bm25 import LexicalIndex, TableStats
from .store import NativeStore

try:  # the reference subclasses langchain_core's VectorStore (postgres_vectorstore.py:16-18)
    from langchain_core.documents import Document  # type: ignore
    from langchain_core.vectorstores import VectorStore as _VectorStoreBase  # type: ignore
except Exception:  # langchain is not installed in the build image: same surface, local base

    class Document:  # minimal stand-in for langchain_core.documents.Document
        def __init__(self, page_content: str, metadata: Optional[Dict[str, Any]] = None, **kwargs: Any):
            self.page_content = page_content
            self.metadata = metadata if metadata is not None else {}
            for k, v in kwargs.items():
                setattr(self, k, v)

        def __repr__(self) -> str:
            return f"Document(page_content={self.page_content!r}, metadata={self.metadata!r})"

        def __eq__(self, other) -> bool:
            return (isinstance(other, Document) and self.page_content == other.page_content
                    and self.metadata == other.metadata)

    class _VectorStoreBase:  # noqa: D401 - stand-in base
        pass


DISTANCE_OPS = {"cosine": "<=>", "l2": "<->", "inner_product": "<#>"}  # postgres_vectorstore.py:74-78
_DOC_FIELDS = ("resource_hash", "display_name", "source_type", "url")    # postgres_vectorstore.py:347-354


class _Collection:
    """GPU-resident state of one collection, shared by every store object that names it."""

    def __init__(self, name: str, metric: str, device: int, storage_dtype: str, bm25_index: bool,
                 table: Optional[TableStats] = None):
        self.name, self.metric, self.device, self.storage_dtype = name, metric, device, storage_dtype
        self.native: Optional[NativeStore] = None
        self.texts: List[str] = []
        self.metadatas: List[Dict[str, Any]] = []
        self.document_ids: List[Any] = []
        self.chunk_index: List[int] = []
        self.live: List[bool] = []
        self.by_chunk_id: Dict[str, List[int]] = {}
        self.by_doc_chunk: Dict[Tuple[Any, int], int] = {}
        self.by_document: Dict[Any, List[int]] = {}
        self.documents: Dict[Any, Dict[str, Any]] = {}  # document-level metadata + is_deleted
        self.lexical: Optional[LexicalIndex] = LexicalIndex(device, table=table) if bm25_index else None
        # metadata equality filters: per key, value text -> row ids, extended incrementally as rows are added
        self.filter_index: Dict[str, Dict[str, Any]] = {}
        # packed device bitmasks per (filter, include_deleted), valid for (rows, docs_epoch)
        self.mask_cache: Dict[Tuple, Tuple[int, int, Any]] = {}
        self.docs_epoch = 0             # bumped when a document's is_deleted flag changes
        self.lock = threading.RLock()

    def ensure_native(self, dim: int) -> NativeStore:
        if self.native is None:
            self.native = NativeStore(dim, self.metric, self.storage_dtype, self.device)
        elif self.native.dim != dim:
            raise ValueError(f"expected {self.native.dim} dimensions, not {dim}")  # pgvector's error text
        return self.native


_REGISTRY: Dict[Tuple[str, str, int], _Collection] = {}
_TABLES: Dict[Tuple[str, int], TableStats] = {}      # BM25 statistics span the collections of one database
_REGISTRY_LOCK = threading.Lock()


def _database_key(pg_config: Optional[Dict[str, Any]]) -> str:
    """Which ``document_chunks`` table a store object talks to: the reference opens ``psycopg2.connect(**pg_config)``
    (postgres_vectorstore.py:94-98), so two configs naming the same server and database share the table."""
    if not pg_config:
        return "default"
    return "{}:{}/{}".format(pg_config.get("host", ""), pg_config.get("port", ""), pg_config.get("dbname", ""))


class B200VectorStore(_VectorStoreBase):
    def __init__(
        self,
        pg_config: Optional[Dict[str, Any]] = None,
        embedding_function: Any = None,
        collection_name: str = "default",
        distance_metric: str = "cosine",
        *,
        connection: Any = None,
        device: int = 0,
        storage_dtype: str = "f32",
        bm25_index: bool = True,
    ):
        """Same positional arguments as PostgresVectorStore.__init__ (:47-56).  ``pg_config`` and
        ``connection`` are accepted and ignored (no database on this path).  Extra keyword-only
        arguments choose the GPU, the storage dtype ('f32' | 'bf16') and whether the BM25 index
        exists (the reference creates it in init.sql:297-300; without it hybrid_search raises)."""
        self._pg_config = pg_config
        self._embedding_function = embedding_function
        self._collection_name = collection_name
        self._distance_metric = distance_metric
        self._external_connection = connection
        self._distance_ops = dict(DISTANCE_OPS)
        if distance_metric not in self._distance_ops:
            raise ValueError(f"distance_metric must be one of {list(self._distance_ops.keys())}")
        self._distance_op = self._distance_ops[distance_metric]
        db = _database_key(pg_config)
        key = (db, collection_name, int(device))
        with _REGISTRY_LOCK:
            coll = _REGISTRY.get(key)
            if coll is None:
                table = _TABLES.setdefault((db, int(device)), TableStats())
                coll = _Collection(collection_name, distance_metric, int(device), storage_dtype, bm25_index, table)
                _REGISTRY[key] = coll
            elif coll.metric != distance_metric:
                raise ValueError(
                    f"collection {collection_name!r} was created with distance_metric={coll.metric!r}; "
                    "the metric is fixed per collection (as the operator class is in init.sql:282)")
        self._coll = coll

    # ---- registry helpers (no reference counterpart: the table outlives the Python object) ------
    @classmethod
    def drop_collection(cls, collection_name: str, device: int = 0, pg_config: Optional[Dict[str, Any]] = None) -> None:
        with _REGISTRY_LOCK:
            coll = _REGISTRY.pop((_database_key(pg_config), collection_name, int(device)), None)
        if coll is not None:
            if coll.lexical is not None:
                coll.lexical.detach()
            if coll.native is not None:
                coll.native.close()

    @property
    def embeddings(self):
        return self._embedding_function

    @property
    def native(self) -> Optional[NativeStore]:
        return self._coll.native

    # ---- documents table stand-in ------------------------------------------------------------------
    def register_document(self, document_id: Any, *, is_deleted: bool = False, **fields: Any) -> None:
        """Document-level columns the reference joins in (documents d: resource_hash, display_name,
        source_type, url, is_deleted; postgres_vectorstore.py:323-328, 304-308)."""
        with self._coll.lock:
            rec = self._coll.documents.setdefault(document_id, {})
            rec.update({k: v for k, v in fields.items() if k in _DOC_FIELDS})
            if rec.get("is_deleted", False) != bool(is_deleted):
                self._coll.docs_epoch += 1
            rec["is_deleted"] = bool(is_deleted)

    # ---- add ------------------------------------------------------------------------------------------
    def add_texts(self, texts: Iterable[str], metadatas: Optional[List[Dict[str, Any]]] = None, *,
                  ids: Optional[List[str]] = None, **kwargs: Any) -> List[str]:
        """postgres_vectorstore.py:105-186."""
        texts_list = list(texts)
        if not texts_list:
            return []
        if ids is None:
            ids = [str(uuid.uuid4()) for _ in texts_list]
        if metadatas is None:
            metadatas = [{} for _ in texts_list]
        for meta in metadatas:
            meta["collection"] = self._collection_name
        document_id = kwargs.get("document_id")
        ef = self._embedding_function

        def embed_and_append(coll):
            if hasattr(ef, "embed_documents_into"):
                # B200Embeddings: encoder forward -> fused pool+normalise kernel writes the rows
                return ef.embed_documents_into(texts_list, coll)
            embeddings = ef.embed_documents(texts_list)
            arr = np.asarray(embeddings, dtype=np.float32)
            if arr.ndim != 2 or arr.shape[0] != len(texts_list):
                raise ValueError("embed_documents must return one vector per text")
            return coll.ensure_native(arr.shape[1]).append(arr)

        return self._insert_rows(texts_list, metadatas, ids, document_id, embed_and_append)

    def _insert_rows(self, texts_list: List[str], metadatas: List[Dict[str, Any]], ids: List[str], document_id: Any,
                     append_rows) -> List[str]:
        """Bookkeeping shared by add_texts / add_embedded_texts: upsert on (document_id, chunk_index) -- the
        replaced rows become tombstones (:173-176) --, chunk_id stamping (:157), host-side text / metadata /
        lexical index.  ``append_rows(coll)`` puts the embeddings into the native store and returns the first row."""
        coll = self._coll
        with coll.lock:
            replaced = []
            if document_id is not None:
                for i in range(len(texts_list)):
                    old = coll.by_doc_chunk.get((document_id, i))
                    if old is not None and coll.live[old]:
                        replaced.append(old)
            first = append_rows(coll)
            if replaced:
                self._tombstone(replaced)
            for i, (text, metadata, chunk_id) in enumerate(zip(texts_list, metadatas, ids)):
                metadata["chunk_id"] = chunk_id
                row = first + i
                assert row == len(coll.texts)
                coll.texts.append(text)
                coll.metadatas.append(dict(metadata))
                coll.document_ids.append(document_id)
                coll.chunk_index.append(i)
                coll.live.append(True)
                coll.by_chunk_id.setdefault(chunk_id, []).append(row)
                if document_id is not None:
                    coll.by_doc_chunk[(document_id, i)] = row
                    coll.by_document.setdefault(document_id, []).append(row)
            if coll.lexical is not None:
                coll.lexical.add_texts(texts_list)
        return ids

    def add_embedded_texts(self, texts: Iterable[str], embeddings: Any, metadatas: Optional[List[Dict[str, Any]]] = None,
                           *, ids: Optional[List[str]] = None, document_id: Any = None) -> List[str]:
        """Insert chunks whose embeddings already exist -- the VectorStoreManager path, which calls
        ``embed_documents`` itself and then INSERTs (manager.py:373, 397-422).  ``embeddings``: [n, D] torch
        CUDA tensor (stays on the device), numpy array or list of vectors, one per text.  Ids, metadata
        stamping and the upsert on (document_id, chunk_index) are those of ``add_texts``."""
        texts_list = list(texts)
        if not texts_list:
            return []
        if ids is None:
            ids = [str(uuid.uuid4()) for _ in texts_list]
        if metadatas is None:
            metadatas = [{} for _ in texts_list]
        if len(metadatas) != len(texts_list) or len(ids) != len(texts_list):
            raise ValueError("texts, metadatas and ids must have the same length")
        if hasattr(embeddings, "is_cuda"):
            emb = embeddings if embeddings.dim() == 2 else embeddings.reshape(len(texts_list), -1)
            n_emb, dim = int(emb.shape[0]), int(emb.shape[1])
        else:
            emb = np.asarray(embeddings, dtype=np.float32)
            if emb.ndim != 2:
                raise ValueError("embeddings must be a [n, D] matrix")
            n_emb, dim = emb.shape
        if n_emb != len(texts_list):
            raise ValueError("embeddings must hold one vector per text")
        for meta in metadatas:
            meta["collection"] = self._collection_name
        return self._insert_rows(texts_list, metadatas, ids, document_id, lambda coll: coll.ensure_native(dim).append(emb))

    def add_documents(self, documents: List[Document], **kwargs: Any) -> List[str]:
        """postgres_vectorstore.py:188-205."""
        texts = [doc.page_content for doc in documents]
        metadatas = [doc.metadata for doc in documents]
        return self.add_texts(texts, metadatas=metadatas, **kwargs)

    # ---- search -----------------------------------------------------------------------------------------
    def similarity_search(self, query: str, k: int = 4, **kwargs: Any) -> List[Document]:
        docs_and_scores = self.similarity_search_with_score(query, k=k, **kwargs)
        return [doc for doc, _ in docs_and_scores]

    def similarity_search_with_score(self, query: str, k: int = 4, **kwargs: Any) -> List[Tuple[Document, float]]:
        query_embedding = self._embedding_function.embed_query(query)
        return self.similarity_search_by_vector_with_score(query_embedding, k=k, **kwargs)

    def similarity_search_by_vector(self, embedding: List[float], k: int = 4, **kwargs: Any) -> List[Document]:
        docs_and_scores = self.similarity_search_by_vector_with_score(embedding, k=k, **kwargs)
        return [doc for doc, _ in docs_and_scores]

    def similarity_search_by_vector_with_score(self, embedding: List[float], k: int = 4,
                                               **kwargs: Any) -> List[Tuple[Document, float]]:
        """postgres_vectorstore.py:272-364: ascending distance; score = 1 - distance for cosine,
        the raw distance for l2 and the negative inner product for inner_product (:361)."""
        metadata_filter = kwargs.get("filter", {}) or {}
        include_deleted = kwargs.get("include_deleted", False)
        coll = self._coll
        with coll.lock:
            if coll.native is None or k <= 0:
                return []
            mask = self._where_mask(metadata_filter, include_deleted)
            scores, ids = coll.native.search(np.asarray(embedding, dtype=np.float32), k, filter_mask=mask)
            return self._rows_to_results(ids[0], scores[0])

    def hybrid_search(self, query: str, k: int = 4, *, semantic_weight: float = 0.7, bm25_weight: float = 0.3,
                      **kwargs: Any) -> List[Tuple[Document, float]]:
        """postgres_vectorstore.py:366-491: combined = (1 - distance)*semantic_weight +
        COALESCE(bm25, 0)*bm25_weight, best first; RuntimeError when there is no BM25 index
        (:415-418); zero rows fall back to similarity_search_with_score (:468-469)."""
        query_embedding = self._embedding_function.embed_query(query)
        metadata_filter = kwargs.get("filter", {}) or {}
        include_deleted = kwargs.get("include_deleted", False)
        coll = self._coll
        if coll.lexical is None:
            raise RuntimeError("Hybrid search requires pg_textsearch BM25 index on document_chunks; none found.")
        with coll.lock:
            results: List[Tuple[Document, float]] = []
            if coll.native is not None and k > 0:
                mask = self._where_mask(metadata_filter, include_deleted)
                # posting lists of the query terms go straight to the kernel: no per-row BM25 vector
                scores, ids = coll.native.hybrid_search_terms(coll.lexical, [query], np.asarray(query_embedding, dtype=np.float32),
                                                              k, semantic_weight, bm25_weight, filter_mask=mask)
                results = self._rows_to_results(ids[0], scores[0])
        if not results:
            return self.similarity_search_with_score(query, k=k, **kwargs)
        return results

    # ---- delete / count / from_texts --------------------------------------------------------------------
    def delete(self, ids: Optional[List[str]] = None, **kwargs: Any) -> Optional[bool]:
        """postgres_vectorstore.py:493-535."""
        document_id = kwargs.get("document_id")
        if ids is None and document_id is None:
            return False
        coll = self._coll
        with coll.lock:
            rows: List[int] = []
            if document_id is not None:
                rows = [r for r in coll.by_document.get(document_id, []) if coll.live[r]]
            elif ids:
                for chunk_id in ids:
                    rows.extend(r for r in coll.by_chunk_id.get(chunk_id, []) if coll.live[r])
            if rows:
                self._tombstone(rows)
        return True

    def count(self) -> int:
        """postgres_vectorstore.py:570-585."""
        coll = self._coll
        with coll.lock:
            return 0 if coll.native is None else coll.native.count()

    @classmethod
    def from_texts(cls: Type["B200VectorStore"], texts: List[str], embedding: Any,
                   metadatas: Optional[List[Dict[str, Any]]] = None, **kwargs: Any) -> "B200VectorStore":
        """postgres_vectorstore.py:537-568 (``pg_config`` is optional here)."""
        pg_config = kwargs.pop("pg_config", None)
        collection_name = kwargs.pop("collection_name", "default")
        distance_metric = kwargs.pop("distance_metric", "cosine")
        ctor = {k: kwargs.pop(k) for k in ("device", "storage_dtype", "bm25_index") if k in kwargs}
        store = cls(pg_config=pg_config, embedding_function=embedding, collection_name=collection_name,
                    distance_metric=distance_metric, **ctor)
        store.add_texts(texts, metadatas=metadatas, **kwargs)
        return store

    # ---- internals ---------------------------------------------------------------------------------------
    def _tombstone(self, rows: List[int]) -> None:
        coll = self._coll
        coll.native.delete_rows(rows)
        for r in rows:
            coll.live[r] = False
        if coll.lexical is not None:
            coll.lexical.delete_rows(rows)

    def _filter_rows(self, key: str, value_text: str) -> List[int]:
        """Row ids whose ``metadata->>key`` equals ``value_text``.  The per-key index is built in one pass
        over the rows and extended in place as rows are added (never rebuilt, never per distinct value)."""
        coll = self._coll
        n = len(coll.metadatas)
        index = coll.filter_index.get(key)
        if index is None:
            index = coll.filter_index[key] = {"n": 0, "rows": {}}
        if index["n"] < n:
            rows = index["rows"]
            for r in range(index["n"], n):
                v = coll.metadatas[r].get(key)
                if v is not None:
                    rows.setdefault(_json_text(v), []).append(r)
            index["n"] = n
        return index["rows"].get(value_text, [])

    def _where_mask(self, metadata_filter: Dict[str, Any], include_deleted: bool):
        """The WHERE clause (:296-310) as a device bitmask, or None when every row passes.
        ``metadata->>'key' = str(value)`` per filter key; documents flagged is_deleted are excluded
        unless include_deleted.  (Rows removed with delete() are tombstoned in the native store.)
        The packed mask stays on the device, keyed by (filter, include_deleted), until rows are added or a
        document's is_deleted flag changes."""
        coll = self._coll
        n = len(coll.texts)
        gone = [] if include_deleted else [d for d, rec in coll.documents.items() if rec.get("is_deleted")]
        if not metadata_filter and not gone:
            return None
        cache_key = (tuple(sorted((str(k), str(v)) for k, v in metadata_filter.items())), bool(include_deleted))
        hit = coll.mask_cache.get(cache_key)
        if hit is not None and hit[0] == n and hit[1] == coll.docs_epoch:
            return hit[2]
        keep: Optional[np.ndarray] = None
        for key, value in metadata_filter.items():
            sel = np.zeros(n, dtype=bool)
            sel[np.asarray(self._filter_rows(key, str(value)), dtype=np.int64)] = True
            keep = sel if keep is None else (keep & sel)
        if gone:
            sel = np.ones(n, dtype=bool)
            for d in gone:
                sel[np.asarray(coll.by_document.get(d, []), dtype=np.int64)] = False
            keep = sel if keep is None else (keep & sel)
        pad = (-n) % 32
        bits = np.concatenate([keep, np.zeros(pad, dtype=bool)]) if pad else keep
        words = np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
        words = np.concatenate([words, np.zeros(1, dtype=np.uint32)])
        mask = _upload_mask_words(words, coll.device)
        if len(coll.mask_cache) >= 64:
            coll.mask_cache.clear()
        coll.mask_cache[cache_key] = (n, coll.docs_epoch, mask)
        return mask

    def _rows_to_results(self, ids: np.ndarray, scores: np.ndarray) -> List[Tuple[Document, float]]:
        coll = self._coll
        results: List[Tuple[Document, float]] = []
        for row, score in zip(ids.tolist(), scores.tolist()):
            if row < 0:
                break
            metadata = dict(coll.metadatas[row] or {})
            rec = coll.documents.get(coll.document_ids[row])
            if rec:
                for f in _DOC_FIELDS:
                    if rec.get(f):
                        metadata[f] = rec[f]
            results.append((Document(page_content=coll.texts[row], metadata=metadata), float(score)))
        return results


def _upload_mask_words(words: np.ndarray, device: int):
    """uint32 bitmask words -> int32 CUDA tensor on the collection's GPU."""
    import torch
    return torch.from_numpy(words.view(np.int32).copy()).to(torch.device("cuda", device))


def _json_text(v: Any) -> str:
    """What ``metadata->>'key'`` yields for a JSON value: strings as they are, everything else as JSON text
    (booleans lower-case, nested objects / arrays in jsonb's ``{"k": v}`` spelling)."""
    if isinstance(v, str):
        return v
    try:
        return json.dumps(v)
    except (TypeError, ValueError):
        return str(v)
