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


class _Shard:
    """One row shard of a collection: a NativeStore in one GPU's HBM, the posting lists of ITS rows (SURVEY 8e:
    "BM25 postings are sharded by the same row blocks"), and the map from its local row ids to collection rows."""

    def __init__(self, device: int, lexical: Optional[LexicalIndex]):
        self.device = int(device)
        self.native: Optional[NativeStore] = None
        self.lexical = lexical
        self.l2g: List[int] = []            # local row id -> collection row id
        self._l2g_dev = None                # the same as an int64 CUDA tensor (rebuilt when rows were added)

    def l2g_device(self):
        import torch
        if self._l2g_dev is None or self._l2g_dev.shape[0] != len(self.l2g):
            self._l2g_dev = torch.tensor(self.l2g, dtype=torch.int64, device=torch.device("cuda", self.device))
        return self._l2g_dev


class _Collection:
    """GPU-resident state of one collection, shared by every store object that names it.  The rows live in one
    shard per device (``devices``); text, metadata and the id maps stay on the host, indexed by collection row."""

    def __init__(self, name: str, metric: str, devices: List[int], storage_dtype: str, bm25_index: bool,
                 table: Optional[TableStats] = None):
        self.name, self.metric, self.devices, self.storage_dtype = name, metric, list(devices), storage_dtype
        self.device = self.devices[0]
        self.table = table
        self.has_lexical = bool(bm25_index)
        self.shards: List[_Shard] = [_Shard(d, LexicalIndex(d, table=table) if bm25_index else None) for d in self.devices]
        self.texts: List[str] = []
        self.metadatas: List[Dict[str, Any]] = []
        self.document_ids: List[Any] = []
        self.chunk_index: List[int] = []
        self.live: List[bool] = []
        self.row_shard: List[int] = []
        self.row_local: List[int] = []
        self.by_chunk_id: Dict[str, List[int]] = {}
        self.by_doc_chunk: Dict[Tuple[Any, int], int] = {}
        self.by_document: Dict[Any, List[int]] = {}
        self.documents: Dict[Any, Dict[str, Any]] = {}  # document-level metadata + is_deleted
        # metadata equality filters: per key, value text -> row ids, extended incrementally as rows are added
        self.filter_index: Dict[str, Dict[str, Any]] = {}
        # packed device bitmasks per (filter, include_deleted, shard), valid for (rows, docs_epoch)
        self.mask_cache: Dict[Tuple, Tuple[int, int, Any]] = {}
        self.docs_epoch = 0             # bumped when a document's is_deleted flag changes
        self.dim: Optional[int] = None
        self.lock = threading.RLock()

    # single-shard views kept for callers that own one GPU (tests, tools, B200Embeddings.embed_documents_into)
    @property
    def native(self) -> Optional[NativeStore]:
        return self.shards[0].native

    @property
    def lexical(self) -> Optional[LexicalIndex]:
        return self.shards[0].lexical

    def ensure_native(self, dim: int, shard: int = 0) -> NativeStore:
        if self.dim is not None and self.dim != dim:
            raise ValueError(f"expected {self.dim} dimensions, not {dim}")  # pgvector's error text
        self.dim = dim
        sh = self.shards[shard]
        if sh.native is None:
            sh.native = NativeStore(dim, self.metric, self.storage_dtype, sh.device)
        return sh.native

    def least_full_shard(self) -> int:
        return min(range(len(self.shards)), key=lambda i: len(self.shards[i].l2g))

    def close(self) -> None:
        for sh in self.shards:
            if sh.lexical is not None:
                sh.lexical.detach()
            if sh.native is not None:
                sh.native.close()
                sh.native = None


class _ShardTarget:
    """What B200Embeddings.embed_documents_into sees: the shard the rows are going to."""

    def __init__(self, coll: _Collection, shard: int):
        self._coll, self._shard = coll, shard

    def ensure_native(self, dim: int) -> NativeStore:
        return self._coll.ensure_native(dim, self._shard)


_REGISTRY: Dict[Tuple[str, str, Tuple[int, ...]], _Collection] = {}
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
        devices: Optional[List[int]] = None,
        storage_dtype: str = "f32",
        bm25_index: bool = True,
    ):
        """Same positional arguments as PostgresVectorStore.__init__ (:47-56).  ``connection`` is accepted and
        ignored, ``pg_config`` only names the database (no database on this path).  Extra keyword-only arguments
        choose the GPU (``device``) or the GPUs the collection is row-sharded over (``devices``: every added batch
        goes to the least-full shard, every search runs on all shards and the per-shard k-lists are merged on the
        first device), the storage dtype ('f32' | 'bf16') and whether the BM25 index exists (the reference creates
        it in init.sql:297-300; without it hybrid_search raises)."""
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
        devs = tuple(int(d) for d in devices) if devices else (int(device),)
        key = (db, collection_name, devs)
        with _REGISTRY_LOCK:
            coll = _REGISTRY.get(key)
            if coll is None:
                table = _TABLES.setdefault((db, devs[0]), TableStats())
                coll = _Collection(collection_name, distance_metric, list(devs), storage_dtype, bm25_index, table)
                _REGISTRY[key] = coll
            elif coll.metric != distance_metric:
                raise ValueError(
                    f"collection {collection_name!r} was created with distance_metric={coll.metric!r}; "
                    "the metric is fixed per collection (as the operator class is in init.sql:282)")
        self._coll = coll

    # ---- registry helpers (no reference counterpart: the table outlives the Python object) ------
    @classmethod
    def drop_collection(cls, collection_name: str, device: int = 0, pg_config: Optional[Dict[str, Any]] = None,
                        devices: Optional[List[int]] = None) -> None:
        devs = tuple(int(d) for d in devices) if devices else (int(device),)
        with _REGISTRY_LOCK:
            coll = _REGISTRY.pop((_database_key(pg_config), collection_name, devs), None)
        if coll is not None:
            coll.close()

    @property
    def embeddings(self):
        return self._embedding_function

    @property
    def native(self) -> Optional[NativeStore]:
        """The NativeStore of the first (or only) shard."""
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

        def embed():
            if hasattr(ef, "embed_documents_device"):
                return ef.embed_documents_device(texts_list)         # stays in HBM (NVLink copy to other shards' GPUs)
            arr = np.asarray(ef.embed_documents(texts_list), dtype=np.float32)
            if arr.ndim != 2 or arr.shape[0] != len(texts_list):
                raise ValueError("embed_documents must return one vector per text")
            return arr

        fused = None
        if hasattr(ef, "embed_documents_into") and len(self._coll.shards) == 1 \
                and getattr(ef, "device", None) == self._coll.shards[0].device:
            # B200Embeddings on the shard's GPU: encoder forward -> fused pool+normalise kernel writes the rows
            fused = lambda coll: ef.embed_documents_into(texts_list, _ShardTarget(coll, 0))  # noqa: E731
        return self._insert_rows(texts_list, metadatas, ids, document_id, embed, fused)

    def _insert_rows(self, texts_list: List[str], metadatas: List[Dict[str, Any]], ids: List[str], document_id: Any,
                     embed, fused=None) -> List[str]:
        """Bookkeeping shared by add_texts / add_embedded_texts: upsert on (document_id, chunk_index) -- the
        replaced rows become tombstones (:173-176) --, chunk_id stamping (:157), host-side text / metadata /
        lexical index.  ``embed()`` returns the [n, D] embeddings (numpy or a CUDA tensor); ``fused(coll)`` instead
        writes them straight into the single shard and returns the first local row.  A multi-shard collection
        spreads the batch over its shards in contiguous slices, emptiest shard first."""
        coll = self._coll
        n = len(texts_list)
        with coll.lock:
            replaced = []
            if document_id is not None:
                for i in range(n):
                    old = coll.by_doc_chunk.get((document_id, i))
                    if old is not None and coll.live[old]:
                        replaced.append(old)
            slices: List[Tuple[int, int, int, int]] = []           # (shard, first local row, begin, end)
            if fused is not None:
                slices.append((0, fused(coll), 0, n))
            else:
                emb = embed()
                order = sorted(range(len(coll.shards)), key=lambda i: len(coll.shards[i].l2g))
                parts = min(len(order), n)
                for j in range(parts):
                    a, b = j * n // parts, (j + 1) * n // parts
                    if b > a:
                        slices.append((order[j], _append_rows(coll, order[j], emb[a:b]), a, b))
            if replaced:
                self._tombstone(replaced)
            for shard, first_local, a, b in slices:
                sh = coll.shards[shard]
                assert first_local == len(sh.l2g)
                for i in range(a, b):
                    metadata, chunk_id = metadatas[i], ids[i]
                    metadata["chunk_id"] = chunk_id
                    row = len(coll.texts)
                    sh.l2g.append(row)
                    coll.row_shard.append(shard)
                    coll.row_local.append(first_local + (i - a))
                    coll.texts.append(texts_list[i])
                    coll.metadatas.append(dict(metadata))
                    coll.document_ids.append(document_id)
                    coll.chunk_index.append(i)
                    coll.live.append(True)
                    coll.by_chunk_id.setdefault(chunk_id, []).append(row)
                    if document_id is not None:
                        coll.by_doc_chunk[(document_id, i)] = row
                        coll.by_document.setdefault(document_id, []).append(row)
                if sh.lexical is not None:
                    sh.lexical.add_texts(texts_list[a:b])
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
        return self._insert_rows(texts_list, metadatas, ids, document_id, lambda: emb)

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
        query_embedding = self._embed_query(query)
        if hasattr(query_embedding, "is_cuda"):
            # device handoff: the query embedding goes from the pool+normalise kernel to the scan in HBM
            return self.similarity_search_by_vector_device(query_embedding, k=k, **kwargs)
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
            if coll.dim is None or k <= 0:
                return []
            q = _host_vector(embedding)
            if len(coll.shards) == 1:
                mask = self._where_mask(metadata_filter, include_deleted, 0)
                scores, ids = coll.native.search(q, k, filter_mask=mask)
                return self._rows_to_results(ids[0], scores[0])
            scores, ids = self._search_shards(q, k, metadata_filter, include_deleted, None)
            return self._rows_to_results(ids, scores)

    def similarity_search_by_vector_device(self, embedding, k: int = 4, **kwargs: Any) -> List[Tuple[Document, float]]:
        """The same search for a query embedding that is already a CUDA tensor ([D] or [1, D], e.g. from
        ``B200Embeddings.embed_query_device``): the vector never visits the host (the reference serialises it as
        decimal text, :313); only the k (score, row id) pairs come back."""
        metadata_filter = kwargs.get("filter", {}) or {}
        include_deleted = kwargs.get("include_deleted", False)
        coll = self._coll
        with coll.lock:
            if coll.dim is None or k <= 0:
                return []
            if len(coll.shards) > 1:
                scores, ids = self._search_shards(embedding, k, metadata_filter, include_deleted, None)
                return self._rows_to_results(ids, scores)
            mask = self._where_mask(metadata_filter, include_deleted, 0)
            scores, ids = coll.native.search(embedding.reshape(1, -1), k, filter_mask=mask)
            return self._rows_to_results(ids[0].cpu().numpy(), scores[0].cpu().numpy())

    def hybrid_search(self, query: str, k: int = 4, *, semantic_weight: float = 0.7, bm25_weight: float = 0.3,
                      **kwargs: Any) -> List[Tuple[Document, float]]:
        """postgres_vectorstore.py:366-491: combined = (1 - distance)*semantic_weight +
        COALESCE(bm25, 0)*bm25_weight, best first; RuntimeError when there is no BM25 index
        (:415-418); zero rows fall back to similarity_search_with_score (:468-469)."""
        query_embedding = self._embed_query(query)
        metadata_filter = kwargs.get("filter", {}) or {}
        include_deleted = kwargs.get("include_deleted", False)
        coll = self._coll
        if not coll.has_lexical:
            raise RuntimeError("Hybrid search requires pg_textsearch BM25 index on document_chunks; none found.")
        with coll.lock:
            results: List[Tuple[Document, float]] = []
            if coll.dim is not None and k > 0:
                if len(coll.shards) == 1:
                    mask = self._where_mask(metadata_filter, include_deleted, 0)
                    # posting lists of the query terms go straight to the kernel: no per-row BM25 vector
                    q = query_embedding if hasattr(query_embedding, "is_cuda") else _host_vector(query_embedding)
                    scores, ids = coll.native.hybrid_search_terms(coll.lexical, [query], q, k, semantic_weight, bm25_weight,
                                                                  filter_mask=mask)
                    if hasattr(scores, "is_cuda"):
                        scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
                    results = self._rows_to_results(ids[0], scores[0])
                else:
                    scores, ids = self._search_shards(query_embedding, k, metadata_filter, include_deleted,
                                                      (query, float(semantic_weight), float(bm25_weight)))
                    results = self._rows_to_results(ids, scores)
        if not results:
            return self.similarity_search_with_score(query, k=k, **kwargs)
        return results

    def _embed_query(self, query: str):
        """embed_query, kept on the device when the embedding function can (B200Embeddings.embed_query_device)."""
        ef = self._embedding_function
        if hasattr(ef, "embed_query_device") and getattr(ef, "device", None) == self._coll.device:
            return ef.embed_query_device(query)
        return ef.embed_query(query)

    def _search_shards(self, embedding, k: int, metadata_filter: Dict[str, Any], include_deleted: bool, hybrid):
        """One query over every shard: the local exact top-k of each GPU is enqueued without waiting for the
        others (device tensors in and out), local row ids are mapped to collection rows on the shard's GPU, the
        k-lists travel to the first device over NVLink and are merged there (archi_merge_topk)."""
        import torch
        from .store import merge_topk
        coll = self._coll
        dev0 = torch.device("cuda", coll.devices[0])
        if hasattr(embedding, "is_cuda"):
            q0 = embedding.reshape(1, -1).to(torch.float32)
        else:
            q0 = torch.from_numpy(_host_vector(embedding)).reshape(1, -1)
        parts_s, parts_i = [], []
        larger = True if hybrid is not None else (coll.metric == "cosine")
        for si, sh in enumerate(coll.shards):
            if sh.native is None or not sh.l2g:
                continue
            dev = torch.device("cuda", sh.device)
            with torch.cuda.device(dev):
                q = q0.to(dev, non_blocking=True)
                mask = self._where_mask(metadata_filter, include_deleted, si)
                if hybrid is not None:
                    text, ws, wb = hybrid
                    sc, ids = sh.native.hybrid_search_terms(sh.lexical, [text], q, k, ws, wb, filter_mask=mask)
                else:
                    sc, ids = sh.native.search(q, k, filter_mask=mask)
                gids = torch.where(ids >= 0, sh.l2g_device()[ids.clamp(min=0)], ids)
            parts_s.append(sc.to(dev0, non_blocking=True))
            parts_i.append(gids.to(dev0, non_blocking=True))
        if not parts_s:
            return np.empty(0, np.float32), np.empty(0, np.int64)
        with torch.cuda.device(dev0):
            if len(parts_s) == 1:
                ms, mi = parts_s[0], parts_i[0]
            else:
                ms, mi = merge_topk(torch.stack(parts_s), torch.stack(parts_i), larger)
            return ms[0].cpu().numpy(), mi[0].cpu().numpy()

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
            return sum(sh.native.count() for sh in coll.shards if sh.native is not None)

    @classmethod
    def from_texts(cls: Type["B200VectorStore"], texts: List[str], embedding: Any,
                   metadatas: Optional[List[Dict[str, Any]]] = None, **kwargs: Any) -> "B200VectorStore":
        """postgres_vectorstore.py:537-568 (``pg_config`` is optional here)."""
        pg_config = kwargs.pop("pg_config", None)
        collection_name = kwargs.pop("collection_name", "default")
        distance_metric = kwargs.pop("distance_metric", "cosine")
        ctor = {k: kwargs.pop(k) for k in ("device", "devices", "storage_dtype", "bm25_index") if k in kwargs}
        store = cls(pg_config=pg_config, embedding_function=embedding, collection_name=collection_name,
                    distance_metric=distance_metric, **ctor)
        store.add_texts(texts, metadatas=metadatas, **kwargs)
        return store

    # ---- snapshot / restore / import -----------------------------------------------------------------
    def save(self, directory: str) -> None:
        """Snapshot of the whole collection: one ``shard<i>.bin`` per GPU shard (archi_store_save: rows, norms,
        tombstones), the posting-list state of each shard's lexical index, and the host side (texts, metadata,
        document ids, chunk indices, row placement, the documents table stand-in).  The reference's persistence is the
        table itself (init.sql:256-276); this is what lets a GPU-resident store restart without re-embedding."""
        import gzip
        import os
        coll = self._coll
        os.makedirs(directory, exist_ok=True)
        with coll.lock:
            for i, sh in enumerate(coll.shards):
                if sh.native is not None:
                    sh.native.save(os.path.join(directory, f"shard{i}.bin"))
                if sh.lexical is not None:
                    sh.lexical.save(os.path.join(directory, f"lexical{i}.npz"))
            manifest = {"format": "archi_b200 collection snapshot v1", "collection": self._collection_name,
                        "metric": coll.metric, "storage_dtype": coll.storage_dtype, "dim": coll.dim,
                        "n_shards": len(coll.shards), "rows": len(coll.texts), "bm25_index": coll.has_lexical,
                        "shard_rows": [len(sh.l2g) for sh in coll.shards]}
            with open(os.path.join(directory, "manifest.json"), "w") as f:
                json.dump(manifest, f)
            with gzip.open(os.path.join(directory, "rows.jsonl.gz"), "wt", encoding="utf-8") as f:
                for r in range(len(coll.texts)):
                    f.write(json.dumps([coll.texts[r], coll.metadatas[r], coll.document_ids[r], coll.chunk_index[r],
                                        coll.live[r], coll.row_shard[r], coll.row_local[r]], ensure_ascii=False) + "\n")
            with open(os.path.join(directory, "documents.json"), "w") as f:
                json.dump([[k, v] for k, v in coll.documents.items()], f)

    @classmethod
    def load(cls, directory: str, embedding_function: Any, *, pg_config: Optional[Dict[str, Any]] = None,
             collection_name: Optional[str] = None, device: int = 0, devices: Optional[List[int]] = None) -> "B200VectorStore":
        """Restore a snapshot written by ``save`` onto the same number of GPUs (any device ids)."""
        import gzip
        import os
        with open(os.path.join(directory, "manifest.json")) as f:
            m = json.load(f)
        devs = [int(d) for d in devices] if devices else [int(device)]
        if len(devs) != m["n_shards"]:
            raise ValueError(f"the snapshot has {m['n_shards']} shard(s); {len(devs)} device(s) were given")
        name = collection_name or m["collection"]
        cls.drop_collection(name, pg_config=pg_config, devices=devs)
        store = cls(pg_config, embedding_function, collection_name=name, distance_metric=m["metric"], devices=devs,
                    storage_dtype=m["storage_dtype"], bm25_index=m["bm25_index"])
        coll = store._coll
        with coll.lock:
            coll.dim = m["dim"]
            for i, sh in enumerate(coll.shards):
                path = os.path.join(directory, f"shard{i}.bin")
                if os.path.exists(path):
                    sh.native = NativeStore.load(path, sh.device)
                lex = os.path.join(directory, f"lexical{i}.npz")
                if sh.lexical is not None and os.path.exists(lex):
                    sh.lexical.load(lex)
                sh.l2g = [0] * m["shard_rows"][i]
            with gzip.open(os.path.join(directory, "rows.jsonl.gz"), "rt", encoding="utf-8") as f:
                for r, line in enumerate(f):
                    text, metadata, document_id, chunk_index, live, shard, local = json.loads(line)
                    coll.texts.append(text)
                    coll.metadatas.append(metadata)
                    coll.document_ids.append(document_id)
                    coll.chunk_index.append(chunk_index)
                    coll.live.append(live)
                    coll.row_shard.append(shard)
                    coll.row_local.append(local)
                    coll.shards[shard].l2g[local] = r
                    if isinstance(metadata, dict) and "chunk_id" in metadata:
                        coll.by_chunk_id.setdefault(metadata["chunk_id"], []).append(r)
                    if document_id is not None:
                        coll.by_doc_chunk[(document_id, chunk_index)] = r
                        coll.by_document.setdefault(document_id, []).append(r)
            with open(os.path.join(directory, "documents.json")) as f:
                coll.documents = {(tuple(k) if isinstance(k, list) else k): v for k, v in json.load(f)}
        return store

    def import_pgvector_rows(self, rows: Iterable[Tuple[Any, int, str, str, Any]], batch: int = 8192) -> int:
        """Rebuild from the reference's table: rows of
        ``SELECT document_id, chunk_index, chunk_text, embedding::text, metadata FROM document_chunks`` --
        the embedding in pgvector's text form ``[v1,v2,...]`` (what the reference itself sends, :179,:313), metadata
        as a dict or JSON text.  Stored metadata is kept as it is (chunk_id, collection); rows of other collections
        are skipped like the reference's WHERE clause does (:296).  Returns the number of rows imported."""
        n = 0
        pending: List[Tuple[Any, int, str, np.ndarray, Dict[str, Any]]] = []

        def flush():
            if not pending:
                return
            emb = np.stack([p[3] for p in pending])
            coll = self._coll
            with coll.lock:
                order = sorted(range(len(coll.shards)), key=lambda i: len(coll.shards[i].l2g))
                shard = order[0]
                sh = coll.shards[shard]
                first_local = _append_rows(coll, shard, emb)
                for i, (document_id, chunk_index, text, _, metadata) in enumerate(pending):
                    old = coll.by_doc_chunk.get((document_id, chunk_index)) if document_id is not None else None
                    if old is not None and coll.live[old]:
                        self._tombstone([old])
                    row = len(coll.texts)
                    sh.l2g.append(row)
                    coll.row_shard.append(shard)
                    coll.row_local.append(first_local + i)
                    coll.texts.append(text)
                    coll.metadatas.append(metadata)
                    coll.document_ids.append(document_id)
                    coll.chunk_index.append(chunk_index)
                    coll.live.append(True)
                    if "chunk_id" in metadata:
                        coll.by_chunk_id.setdefault(metadata["chunk_id"], []).append(row)
                    if document_id is not None:
                        coll.by_doc_chunk[(document_id, chunk_index)] = row
                        coll.by_document.setdefault(document_id, []).append(row)
                if sh.lexical is not None:
                    sh.lexical.add_texts([p[2] for p in pending])
            pending.clear()

        for document_id, chunk_index, text, embedding_text, metadata in rows:
            if isinstance(metadata, str):
                metadata = json.loads(metadata)
            metadata = dict(metadata or {})
            if metadata.get("collection") not in (None, self._collection_name):
                continue
            pending.append((document_id, int(chunk_index), text, parse_pgvector_text(embedding_text), metadata))
            n += 1
            if len(pending) >= batch:
                flush()
        flush()
        return n

    # ---- internals ---------------------------------------------------------------------------------------
    def _tombstone(self, rows: List[int]) -> None:
        coll = self._coll
        by_shard: Dict[int, List[int]] = {}
        for r in rows:
            coll.live[r] = False
            by_shard.setdefault(coll.row_shard[r], []).append(coll.row_local[r])
        for si, local in by_shard.items():
            sh = coll.shards[si]
            sh.native.delete_rows(local)
            if sh.lexical is not None:
                sh.lexical.delete_rows(local)

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

    def _where_mask(self, metadata_filter: Dict[str, Any], include_deleted: bool, shard: int = 0):
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
        cache_key = (tuple(sorted((str(k), str(v)) for k, v in metadata_filter.items())), bool(include_deleted), shard)
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
        sh = coll.shards[shard]
        if len(coll.shards) > 1:
            keep = keep[np.asarray(sh.l2g, dtype=np.int64)]          # the shard's rows, in local order
        pad = (-keep.size) % 32
        bits = np.concatenate([keep, np.zeros(pad, dtype=bool)]) if pad else keep
        words = np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
        words = np.concatenate([words, np.zeros(1, dtype=np.uint32)])
        mask = _upload_mask_words(words, sh.device)
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


def parse_pgvector_text(text: str) -> np.ndarray:
    """pgvector's text form ``[v1,v2,...]`` -> float32 vector (the ``::vector`` cast stores float4)."""
    t = text.strip()
    if not (t.startswith("[") and t.endswith("]")):
        raise ValueError(f"malformed vector literal: {text[:40]!r}")
    body = t[1:-1].strip()
    if not body:
        return np.empty(0, dtype=np.float32)
    return np.asarray(body.split(","), dtype=np.float64).astype(np.float32)


def _host_vector(embedding: Any) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(embedding, dtype=np.float32).reshape(-1))


def _append_rows(coll: _Collection, shard: int, emb: Any) -> int:
    """[n, D] embeddings (numpy, or a CUDA tensor on any GPU of the box) -> the shard's native store."""
    sh = coll.shards[shard]
    if hasattr(emb, "is_cuda"):
        native = coll.ensure_native(int(emb.shape[1]), shard)
        if not emb.is_cuda or emb.device.index != sh.device:
            import torch
            emb = emb.to(torch.device("cuda", sh.device))
        return native.append(emb)
    return coll.ensure_native(int(emb.shape[1]), shard).append(emb)


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
