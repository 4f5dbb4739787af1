"""IngestionDriver -- the caller of the embed + insert half of the hot path.

Mirrors ``VectorStoreManager._add_to_postgres`` (src/data_manager/vectorstore/manager.py:253-457):
the same per-file status machine (``embedding`` -> ``embedded`` | ``failed`` with the error text), the
same chunking (CharacterTextSplitter("\\n\\n", chunk_size, chunk_overlap), manager.py:75-78,292), the
same chunk metadata (``chunk_index, filename, resource_hash, collection`` over the file-level metadata,
manager.py:314-320), NUL bytes stripped, blank chunks skipped, failures isolated per file (the
reference's SAVEPOINT / ROLLBACK TO SAVEPOINT), a commit every ``commit_batch_size`` = 25 files.

What changes is how the GPU is fed.  The reference embeds one file at a time (serial
``embed_documents(chunks)`` per file, manager.py:362-373: 30-50 s per file on its CPU encoder).  Here
the chunks of a whole commit group are embedded together: they are ordered by length so that every
encoder batch pads to a similar sequence length, embedded in ONE call (the embeddings stay on the GPU
when the embedding function offers ``embed_documents_device``), put back in file order and appended file
by file with ``B200VectorStore.add_embedded_texts``.  If the group call fails, the group is retried
file by file so that one bad file cannot fail its neighbours.

Loaders for formats that need third-party parsers (pdf, html) are left to the ``loader`` callable the
deployment passes in.
"""
from __future__ import annotations

import logging
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger(__name__)

# suffixes the reference reads as plain text (loader_utils.py:27-28, plus .py through PythonLoader :29-30)
TEXT_SUFFIXES = {".txt", ".c", ".sh", ".h", ".php", ".yaml", ".yml", ".json", ".csv", ".tsv", ".log", ".rst", ".md", ".py"}


def split_text(text: str, chunk_size: int = 1000, chunk_overlap: int = 0, separator: str = "\n\n") -> List[str]:
    """CharacterTextSplitter semantics [external: langchain-text-splitters 1.0.0]: cut at the
    separator, then pack consecutive pieces into chunks of at most ``chunk_size`` characters (joined by
    the separator, stripped); a single piece longer than ``chunk_size`` becomes a chunk of its own (the
    upstream splitter only warns); ``chunk_overlap`` characters' worth of trailing pieces are carried
    into the next chunk."""
    pieces = [p for p in text.split(separator)] if separator else list(text)
    pieces = [p for p in pieces if p != ""]
    sep = len(separator)
    out: List[str] = []
    window: List[str] = []          # pieces of the chunk being built
    size = 0                        # len(separator.join(window))

    def flush() -> None:
        chunk = separator.join(window).strip()
        if chunk:
            out.append(chunk)

    for piece in pieces:
        grown = size + len(piece) + (sep if window else 0)
        if window and grown > chunk_size:
            flush()
            # keep a tail of at most chunk_overlap characters that still leaves room for the new piece
            while window and (size > chunk_overlap or size + len(piece) + sep > chunk_size):
                size -= len(window[0]) + (sep if len(window) > 1 else 0)
                window.pop(0)
        window.append(piece)
        size += len(piece) + (sep if len(window) > 1 else 0)
    if window:
        flush()
    return out


def default_loader(file_path: str) -> Optional[str]:
    """Text of a plain-text file, or None for formats that need a parser (the reference's
    ``select_loader`` returns None -> 'Unsupported file format', manager.py:280-282)."""
    path = Path(file_path)
    if path.suffix.lower() not in TEXT_SUFFIXES:
        return None
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        return f.read()


@dataclass
class IngestReport:
    embedded: List[str] = field(default_factory=list)          # file hashes now 'embedded'
    failed: Dict[str, str] = field(default_factory=dict)       # file hash -> error text
    chunks: int = 0                                            # rows appended
    commits: int = 0
    embed_calls: int = 0                                       # calls into the embedding function
    group_retries: int = 0                                     # commit groups re-run file by file


class _NullCatalog:
    """Stand-in when no catalog is given: statuses are only reported."""

    def update_ingestion_status(self, filehash: str, status: str, error: Optional[str] = None) -> None:
        pass

    def get_document_id(self, filehash: str) -> Any:
        return filehash

    def get_metadata_for_hash(self, filehash: str) -> Dict[str, Any]:
        return {}

    def commit(self) -> None:
        pass


class IngestionDriver:
    """``add_files({resource_hash: path})`` with the reference's semantics, batched for the GPU.

    ``store``: a B200VectorStore (anything with ``add_embedded_texts``, ``embeddings`` and
    ``_collection_name``).  ``catalog``: optional object with any of ``update_ingestion_status(hash,
    status, error=None)``, ``get_document_id(hash)``, ``get_metadata_for_hash(hash)``, ``commit()`` (the
    reference's CatalogService surface used on this path)."""

    def __init__(self, store, *, catalog: Any = None, loader: Callable[[str], Optional[str]] = default_loader,
                 chunk_size: int = 1000, chunk_overlap: int = 0, separator: str = "\n\n",
                 commit_batch_size: int = 25, parallel_workers: Optional[int] = None,
                 preprocess: Optional[Callable[[str], str]] = None):
        if not hasattr(store, "add_embedded_texts"):
            raise TypeError("store must provide add_embedded_texts (archi_b200.B200VectorStore)")
        self.store = store
        self.catalog = catalog
        self.loader = loader
        self.chunk_size, self.chunk_overlap, self.separator = int(chunk_size), int(chunk_overlap), separator
        self.commit_batch_size = max(1, int(commit_batch_size))
        default_workers = min(64, (os.cpu_count() or 1) + 4)                     # manager.py:86
        self.parallel_workers = max(1, int(parallel_workers) if parallel_workers is not None else default_workers)
        self.preprocess = preprocess            # e.g. the stemming pass of manager.py:302-304

    # ---- catalog access (every method optional) ------------------------------------------------------
    def _cat(self, name: str):
        fn = getattr(self.catalog, name, None) if self.catalog is not None else None
        return fn if callable(fn) else getattr(_NullCatalog(), name)

    def _status(self, report: IngestReport, filehash: str, status: str, error: Optional[str] = None) -> None:
        update = self._cat("update_ingestion_status")
        if error is not None:
            update(filehash, status, error)
        else:
            update(filehash, status)
        if status == "failed":
            report.failed[filehash] = error or ""
        elif status == "embedded":
            report.embedded.append(filehash)

    def _file_metadata(self, filehash: str) -> Dict[str, str]:
        meta = self._cat("get_metadata_for_hash")(filehash) or {}                # manager.py:505-515
        return {str(k): str(v) for k, v in meta.items() if k is not None and v is not None}

    # ---- step 1: load + split one file (runs in the thread pool) ---------------------------------------
    def _process_file(self, filehash: str, file_path: str) -> Tuple[Optional[Tuple[str, List[str], List[Dict]]], Optional[str]]:
        """-> ((filename, chunks, metadatas), None) or (None, error text)."""
        filename = Path(file_path).name
        try:
            text = self.loader(file_path)
        except Exception as exc:  # noqa: BLE001 - any loader failure fails this file only
            return None, str(exc)
        if text is None:
            return None, f"Unsupported file format: {file_path}"
        file_meta = self._file_metadata(filehash)
        chunks: List[str] = []
        metadatas: List[Dict] = []
        collection = getattr(self.store, "_collection_name", "default")
        for index, chunk in enumerate(split_text(text, self.chunk_size, self.chunk_overlap, self.separator)):
            chunk = chunk.replace("\x00", "")
            if self.preprocess is not None:
                chunk = self.preprocess(chunk)
            if not chunk.strip():
                continue
            chunks.append(chunk)
            meta = dict(file_meta)
            meta["chunk_index"] = index           # index among the splitter's chunks, blanks included
            meta["filename"] = filename
            meta["resource_hash"] = filehash
            meta["collection"] = collection
            metadatas.append(meta)
        if not chunks:
            return None, "No text chunks could be extracted"
        return (filename, chunks, metadatas), None

    # ---- step 2: embed the chunks of several files in one length-ordered pass --------------------------
    def _embed(self, report: IngestReport, texts: Sequence[str]):
        """[len(texts), D] embeddings in input order: a CUDA tensor when the embedding function has
        ``embed_documents_device``, else a float32 numpy array."""
        ef = self.store.embeddings
        order = sorted(range(len(texts)), key=lambda i: len(texts[i]))           # similar lengths share a batch
        ordered = [texts[i] for i in order]
        report.embed_calls += 1
        if hasattr(ef, "embed_documents_device"):
            import torch
            emb = ef.embed_documents_device(ordered)
            inverse = torch.empty(len(order), dtype=torch.long)
            inverse[torch.tensor(order, dtype=torch.long)] = torch.arange(len(order))
            return emb.index_select(0, inverse.to(emb.device))
        emb = np.asarray(ef.embed_documents(ordered), dtype=np.float32)
        if emb.ndim != 2 or emb.shape[0] != len(texts):
            raise ValueError("embed_documents must return one vector per text")
        inverse = np.empty(len(order), dtype=np.int64)
        inverse[np.asarray(order, dtype=np.int64)] = np.arange(len(order))
        return emb[inverse]

    # ---- step 3: one commit group ---------------------------------------------------------------------------
    def _insert_file(self, report: IngestReport, filehash: str, processed, emb) -> None:
        filename, chunks, metadatas = processed
        document_id = self._cat("get_document_id")(filehash)
        if document_id is None:
            logger.warning("No document record found for %s, chunks will have no document_id", filehash)
        try:
            self.store.add_embedded_texts(chunks, emb, metadatas=metadatas, document_id=document_id)
        except Exception as exc:  # noqa: BLE001 - the reference rolls back to the file's savepoint
            logger.error("Failed to store vectors for %s: %s", filename, exc)
            self._status(report, filehash, "failed", str(exc))
            return
        report.chunks += len(chunks)
        self._status(report, filehash, "embedded")

    def _run_group(self, report: IngestReport, group: List[Tuple[str, Any]]) -> None:
        """group: [(filehash, (filename, chunks, metadatas))] in input order."""
        texts: List[str] = [c for _, (_, chunks, _) in group for c in chunks]
        try:
            emb = self._embed(report, texts)
        except Exception as exc:  # noqa: BLE001
            if len(group) == 1:
                logger.error("Failed to embed %s: %s", group[0][1][0], exc)
                self._status(report, group[0][0], "failed", str(exc))
                return
            # one bad file must not fail its neighbours: redo the group file by file
            logger.warning("Embedding a group of %d files failed (%s); retrying file by file", len(group), exc)
            report.group_retries += 1
            for item in group:
                self._run_group(report, [item])
            return
        at = 0
        for filehash, processed in group:
            n = len(processed[1])
            self._insert_file(report, filehash, processed, emb[at:at + n])
            at += n

    # ---- the public entry point --------------------------------------------------------------------------------
    def add_files(self, files_to_add: Dict[str, str]) -> IngestReport:
        report = IngestReport()
        if not files_to_add:
            return report
        items = list(files_to_add.items())
        for filehash, _ in items:                                                 # manager.py:259-261
            self._cat("update_ingestion_status")(filehash, "embedding")
        processed: Dict[str, Any] = {}
        with ThreadPoolExecutor(max_workers=self.parallel_workers) as pool:
            futures = [(filehash, pool.submit(self._process_file, filehash, path)) for filehash, path in items]
            for filehash, fut in futures:
                try:
                    result, error = fut.result()
                except Exception as exc:  # noqa: BLE001
                    result, error = None, str(exc)
                if result is None:
                    self._status(report, filehash, "failed", error)
                else:
                    processed[filehash] = result
        # commit groups follow the INPUT order and count failed files too, as the reference does
        for start in range(0, len(items), self.commit_batch_size):
            group = [(h, processed[h]) for h, _ in items[start:start + self.commit_batch_size] if h in processed]
            if group:
                self._run_group(report, group)
            self._cat("commit")()
            report.commits += 1
        return report
