"""LexicalIndex -- device posting lists for the BM25 half of hybrid_search.

Replaces the pg_textsearch BM25 index the reference requires (``USING bm25`` index probed at
postgres_vectorstore.py:399-418, scored by ``chunk_text <@> to_bm25query(query, index)`` at :433)
[external, parity unpinned: pg_textsearch 0.4.2 is not vendored].  Contract taken from the
reference's own tests/docs: score >= 0, higher is better, rows with no matching term are SQL NULL
(-> COALESCE(.., 0)).  ``sign=-1`` reproduces the literal upstream operator (negated scores).

The host side only tokenises and keeps a term dictionary; posting lists (doc ids, term
frequencies, document lengths) live on the GPU and are scored by archi_bm25_accumulate.
"""
from __future__ import annotations

import ctypes
import math
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _native as N

_TOKEN_RE = re.compile(r"[a-z0-9]+")


def default_tokenize(text: str) -> List[str]:
    """Lower-cased alphanumeric runs (no stemming / stop words; see DESIGN.md assumptions)."""
    return _TOKEN_RE.findall(text.lower())


class LexicalIndex:
    def __init__(self, device: int = 0, k1: float = 1.2, b: float = 0.75, sign: float = 1.0, tokenize=default_tokenize):
        self.device, self.k1, self.b, self.sign, self.tokenize = int(device), float(k1), float(b), float(sign), tokenize
        self._vocab: Dict[str, int] = {}
        self._doc_terms: List[np.ndarray] = []   # per row: sorted unique term ids
        self._doc_tfs: List[np.ndarray] = []     # per row: term frequencies
        self._doc_len: List[int] = []
        self._deleted: set = set()
        self._dirty = True
        # device CSR (built lazily)
        self._post_ptr: Optional[np.ndarray] = None
        self._doc_ids_dev = self._tfs_dev = self._doc_len_dev = None
        self._df: Optional[np.ndarray] = None
        self._n_live = 0
        self._avgdl = 1.0

    def __len__(self) -> int:
        return len(self._doc_len)

    # ---- building ------------------------------------------------------------------------------
    def add_texts(self, texts: Sequence[str]) -> None:
        for t in texts:
            ids = [self._vocab.setdefault(tok, len(self._vocab)) for tok in self.tokenize(t)]
            self.add_token_ids(np.asarray(ids, dtype=np.int64))

    def add_token_ids(self, token_ids: np.ndarray) -> None:
        """One document given as an array of integer term ids (synthetic corpora skip tokenising)."""
        terms, tfs = np.unique(np.asarray(token_ids, dtype=np.int64), return_counts=True)
        self._doc_terms.append(terms)
        self._doc_tfs.append(tfs.astype(np.int32))
        self._doc_len.append(int(token_ids.size))
        self._dirty = True

    def add_token_matrix(self, tokens: np.ndarray) -> None:
        """[n_docs, doc_len] integer term ids, vectorised."""
        for row in np.asarray(tokens):
            self.add_token_ids(row)

    def delete_rows(self, rows) -> None:
        self._deleted.update(int(r) for r in rows)
        self._dirty = True

    def reset(self) -> None:
        self.__init__(self.device, self.k1, self.b, self.sign, self.tokenize)

    def _rebuild(self) -> None:
        import torch
        n = len(self._doc_len)
        live = np.ones(n, dtype=bool)
        if self._deleted:
            live[np.fromiter(self._deleted, dtype=np.int64)] = False
        if n:
            lens = np.asarray([t.size for t in self._doc_terms], dtype=np.int64)
            doc_of = np.repeat(np.arange(n, dtype=np.int64), lens)
            terms = np.concatenate(self._doc_terms) if lens.sum() else np.empty(0, np.int64)
            tfs = np.concatenate(self._doc_tfs) if lens.sum() else np.empty(0, np.int32)
            keep = live[doc_of]
            doc_of, terms, tfs = doc_of[keep], terms[keep], tfs[keep]
        else:
            doc_of, terms, tfs = np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, np.int32)
        n_terms = int(terms.max()) + 1 if terms.size else 0
        order = np.lexsort((doc_of, terms))
        terms, doc_of, tfs = terms[order], doc_of[order], tfs[order]
        self._df = np.bincount(terms, minlength=n_terms).astype(np.int64)
        self._post_ptr = np.concatenate([[0], np.cumsum(self._df)]).astype(np.int64)
        dl = np.asarray(self._doc_len, dtype=np.float32)
        self._n_live = int(live.sum())
        self._avgdl = float(dl[live].mean()) if self._n_live and dl[live].sum() > 0 else 1.0
        dev = torch.device("cuda", self.device)
        self._doc_ids_dev = torch.from_numpy(doc_of.astype(np.int32)).to(dev)
        self._tfs_dev = torch.from_numpy(tfs.astype(np.int32)).to(dev)
        self._doc_len_dev = torch.from_numpy(dl).to(dev)
        self._dirty = False

    # ---- scoring -------------------------------------------------------------------------------
    def query_terms(self, query) -> List[int]:
        if isinstance(query, str):
            return [self._vocab[t] for t in self.tokenize(query) if t in self._vocab]
        return [int(t) for t in np.asarray(query).reshape(-1)]

    def idf(self, term: int) -> float:
        df = int(self._df[term]) if 0 <= term < self._df.size else 0
        return math.log(1.0 + (self._n_live - df + 0.5) / (df + 0.5))

    def score(self, query, out=None):
        """Dense fp32 [rows] BM25 scores of ``query`` on the GPU, 0 where no term matches."""
        import torch
        if self._dirty:
            self._rebuild()
        n = len(self._doc_len)
        dev = torch.device("cuda", self.device)
        if out is None:
            out = torch.zeros(n, dtype=torch.float32, device=dev)
        else:
            out.zero_()
        terms = [t for t in self.query_terms(query) if 0 <= t < self._df.size and self._df[t] > 0]
        if not terms or n == 0:
            return out
        # one C-ABI call for all query-term occurrences (posting ranges are arbitrary, not consecutive)
        starts = np.asarray([self._post_ptr[t] for t in terms], dtype=np.int64)
        ends = np.asarray([self._post_ptr[t + 1] for t in terms], dtype=np.int64)
        idf = np.asarray([self.idf(t) for t in terms], dtype=np.float32)
        stream = ctypes.c_void_p(int(torch.cuda.current_stream(self.device).cuda_stream))
        N.check(N.lib().archi_bm25_accumulate(
            starts.ctypes.data_as(ctypes.c_void_p), ends.ctypes.data_as(ctypes.c_void_p), len(terms),
            idf.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(self._doc_ids_dev.data_ptr()),
            ctypes.c_void_p(self._tfs_dev.data_ptr()), ctypes.c_void_p(self._doc_len_dev.data_ptr()),
            self._avgdl, self.k1, self.b, self.sign, ctypes.c_void_p(out.data_ptr()), stream))
        return out
