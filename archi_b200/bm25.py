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
import os
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _native as N

_TOKEN_RE = re.compile(r"[a-z0-9]+")
_NON_TOKEN_RE = re.compile(r"[^a-z0-9]+")


def default_tokenize(text: str) -> List[str]:
    """Lower-cased alphanumeric runs (no stemming / stop words; see DESIGN.md assumptions)."""
    return _TOKEN_RE.findall(text.lower())


# ---- host helper (archi_b200/hostsrc/text_index.c, plain C): tokenise + per-chunk term counts ------------
_TEXT_LIB = None
_TEXT_LIB_TRIED = False


def _text_lib():
    """libarchi_text.so, or None when it has not been built (the index then runs its Python path:
    same postings, a dictionary instead of hashed term keys)."""
    global _TEXT_LIB, _TEXT_LIB_TRIED
    if not _TEXT_LIB_TRIED:
        _TEXT_LIB_TRIED = True
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libarchi_text.so")
        if os.path.exists(path):
            try:
                lib = ctypes.CDLL(path)
                lib.archi_text_index_batch.restype = ctypes.c_int64
                lib.archi_text_index_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                                       ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
                lib.archi_text_term_key.restype = ctypes.c_uint64
                lib.archi_text_term_key.argtypes = [ctypes.c_char_p, ctypes.c_int64]
                _TEXT_LIB = lib
            except OSError:
                _TEXT_LIB = None
    return _TEXT_LIB


class TableStats:
    """Corpus statistics of one ``document_chunks`` table.  The reference's BM25 index is built on the table
    (init.sql:297-300), so N, df and avgdl come from the chunks of EVERY collection stored in it, while a
    search only returns the rows of one collection (postgres_vectorstore.py:420-430).  Each collection keeps
    its own posting lists (LexicalIndex); the indexes of one database share one TableStats."""

    def __init__(self):
        self.members: List["LexicalIndex"] = []

    def refresh(self) -> None:
        for m in self.members:
            if m._dirty:
                m._rebuild()

    def n_live(self) -> int:
        return sum(m._n_live for m in self.members)

    def avgdl(self) -> float:
        n, total = self.n_live(), sum(m._sum_dl for m in self.members)
        return float(total / n) if n and total > 0 else 1.0

    def df_of_keys(self, keys: np.ndarray) -> np.ndarray:
        out = np.zeros(keys.size, dtype=np.int64)
        for m in self.members:
            out += m._df_of_keys(keys)
        return out


class LexicalIndex:
    """Chunks are kept as (term key, term frequency) pairs.  With the default tokenizer and the host
    helper built, a term key is the 64-bit FNV-1a hash of the token and a whole batch of chunks is
    tokenised and counted in one C call; otherwise (custom ``tokenize``, or integer term ids given
    directly) keys come from a Python dictionary / the ids themselves.  Term ids -- positions in the
    sorted array of distinct keys -- only exist after ``_rebuild``."""

    def __init__(self, device: int = 0, k1: float = 1.2, b: float = 0.75, sign: float = 1.0, tokenize=default_tokenize,
                 table: Optional[TableStats] = None):
        self.device, self.k1, self.b, self.sign, self.tokenize = int(device), float(k1), float(b), float(sign), tokenize
        self.table = table if table is not None else TableStats()
        if self not in self.table.members:
            self.table.members.append(self)
        self._sum_dl = 0.0
        self._fast = tokenize is default_tokenize and _text_lib() is not None
        self._vocab: Dict[str, int] = {}         # Python path only: token -> key
        # per added batch: keys uint64 [pairs], tfs int32 [pairs], pairs per chunk int64 [docs], tokens per chunk int32 [docs]
        self._batches: List[tuple] = []
        self._n_docs = 0
        self._deleted: set = set()
        self._dirty = True
        # device CSR (built lazily)
        self._post_ptr: Optional[np.ndarray] = None
        self._doc_ids_dev = self._tfs_dev = self._doc_len_dev = None
        self._df: Optional[np.ndarray] = None
        self._term_keys = np.empty(0, dtype=np.uint64)
        self._n_live = 0
        self._avgdl = 1.0

    def __len__(self) -> int:
        return self._n_docs

    # ---- building ------------------------------------------------------------------------------
    def add_texts(self, texts: Sequence[str]) -> None:
        texts = list(texts)
        if not texts:
            return
        if self._fast:
            self._add_texts_fast(texts)
            return
        docs = []
        for t in texts:
            ids = [self._vocab.setdefault(tok, len(self._vocab)) for tok in self.tokenize(t)]
            docs.append(np.asarray(ids, dtype=np.int64))
        self._add_id_docs(docs)

    def _add_texts_fast(self, texts: List[str]) -> None:
        # non-ASCII chunks are lower-cased by Python (Unicode rules) and reduced to ASCII tokens first;
        # ASCII chunks go to the helper untouched (it lower-cases A-Z itself)
        blobs = [(t if t.isascii() else _NON_TOKEN_RE.sub(" ", t.lower())).encode("ascii") for t in texts]
        n = len(blobs)
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter(map(len, blobs), dtype=np.int64, count=n), out=offs[1:])
        text = b"".join(blobs)
        cap = int(offs[-1]) // 2 + n + 1                      # a token needs a byte and a separator
        keys = np.empty(cap, dtype=np.uint64)
        tfs = np.empty(cap, dtype=np.int32)
        doc_ptr = np.empty(n + 1, dtype=np.int64)
        doc_len = np.empty(n, dtype=np.int32)
        w = _text_lib().archi_text_index_batch(text, offs.ctypes.data, n, keys.ctypes.data, tfs.ctypes.data, cap,
                                               doc_ptr.ctypes.data, doc_len.ctypes.data)
        if w < 0:
            raise RuntimeError(f"archi_text_index_batch failed with code {w}")
        self._batches.append((keys[:w].copy(), tfs[:w].copy(), np.diff(doc_ptr), doc_len))
        self._n_docs += n
        self._dirty = True

    def _add_id_docs(self, docs: List[np.ndarray]) -> None:
        """Chunks given as arrays of non-negative integer term keys (one entry per token occurrence)."""
        keys, tfs, counts, lens = [], [], [], []
        for ids in docs:
            terms, tf = np.unique(np.asarray(ids, dtype=np.int64), return_counts=True)
            keys.append(terms.astype(np.uint64))
            tfs.append(tf.astype(np.int32))
            counts.append(terms.size)
            lens.append(int(np.asarray(ids).size))
        self._batches.append((np.concatenate(keys) if keys else np.empty(0, np.uint64),
                              np.concatenate(tfs) if tfs else np.empty(0, np.int32),
                              np.asarray(counts, dtype=np.int64), np.asarray(lens, dtype=np.int32)))
        self._n_docs += len(docs)
        self._dirty = True

    def add_token_ids(self, token_ids: np.ndarray) -> None:
        """One document given as an array of integer term ids (synthetic corpora skip tokenising)."""
        self._add_id_docs([np.asarray(token_ids, dtype=np.int64)])

    def add_token_matrix(self, tokens: np.ndarray) -> None:
        """[n_docs, doc_len] integer term ids, reduced to (term, frequency) pairs in one vectorised pass."""
        t = np.sort(np.asarray(tokens, dtype=np.int64), axis=1)
        if t.ndim != 2:
            raise ValueError("tokens must be a [n_docs, doc_len] matrix")
        n, length = t.shape
        if n == 0:
            return
        if length == 0:
            self._add_id_docs([np.empty(0, np.int64)] * n)
            return
        start = np.ones(t.shape, dtype=bool)                  # first occurrence of a term in its (sorted) row
        start[:, 1:] = t[:, 1:] != t[:, :-1]
        at = np.flatnonzero(start.ravel())
        tfs = np.diff(np.append(at, t.size)).astype(np.int32)  # a run never crosses a row: every row starts one
        self._batches.append((t[start].astype(np.uint64), tfs, start.sum(axis=1).astype(np.int64),
                              np.full(n, length, dtype=np.int32)))
        self._n_docs += n
        self._dirty = True

    def delete_rows(self, rows) -> None:
        self._deleted.update(int(r) for r in rows)
        self._dirty = True

    def reset(self) -> None:
        self.__init__(self.device, self.k1, self.b, self.sign, self.tokenize, self.table)

    def save(self, path: str) -> None:
        """The index state (per-chunk (term key, tf) pairs, chunk lengths, tombstones) as one .npz; the device
        posting lists are rebuilt from it on the first search after ``load``."""
        if self._batches:
            keys = np.concatenate([b[0] for b in self._batches])
            tfs = np.concatenate([b[1] for b in self._batches])
            counts = np.concatenate([b[2] for b in self._batches])
            lens = np.concatenate([b[3] for b in self._batches])
        else:
            keys, tfs = np.empty(0, np.uint64), np.empty(0, np.int32)
            counts, lens = np.empty(0, np.int64), np.empty(0, np.int32)
        vocab = np.asarray(sorted(self._vocab.items(), key=lambda kv: kv[1]), dtype=object) if self._vocab else np.empty((0, 2), dtype=object)
        np.savez_compressed(path, keys=keys, tfs=tfs, counts=counts, lens=lens,
                            deleted=np.asarray(sorted(self._deleted), dtype=np.int64),
                            params=np.asarray([self.k1, self.b, self.sign, float(self._fast)]),
                            vocab_tokens=np.asarray([t for t, _ in vocab], dtype=str) if len(vocab) else np.empty(0, dtype=str))

    def load(self, path: str) -> None:
        z = np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=False)
        self._batches = [(z["keys"], z["tfs"], z["counts"], z["lens"])] if z["counts"].size else []
        self._n_docs = int(z["counts"].size)
        self._deleted = set(int(r) for r in z["deleted"])
        self.k1, self.b, self.sign = (float(v) for v in z["params"][:3])
        if bool(z["params"][3]) != self._fast:
            raise RuntimeError("the snapshot's lexical index was built with a different tokeniser path "
                               "(libarchi_text.so present / absent): re-tokenise with add_texts instead")
        self._vocab = {str(t): i for i, t in enumerate(z["vocab_tokens"])}
        self._dirty = True

    def detach(self) -> None:
        """Leave the table statistics (the collection is dropped)."""
        if self in self.table.members:
            self.table.members.remove(self)

    def _df_of_keys(self, keys: np.ndarray) -> np.ndarray:
        """Document frequency, in THIS index, of each term key (0 for unknown keys)."""
        if self._df is None or self._term_keys.size == 0 or keys.size == 0:
            return np.zeros(keys.size, dtype=np.int64)
        idx = np.searchsorted(self._term_keys, keys)
        idx[idx >= self._term_keys.size] = 0
        return np.where(self._term_keys[idx] == keys, self._df[idx], 0).astype(np.int64)

    def _host_csr(self):
        """Posting lists on the host: (term_keys sorted uint64 [T], df int64 [T], post_ptr int64 [T+1],
        doc_ids int32 [P] grouped by term and ascending inside a term, tfs int32 [P], doc_len float32
        [docs], n_live, avgdl).  Deleted chunks keep their row but have no postings."""
        n = self._n_docs
        live = np.ones(n, dtype=bool)
        if self._deleted:
            live[np.fromiter(self._deleted, dtype=np.int64)] = False
        if self._batches:
            keys = np.concatenate([b[0] for b in self._batches])
            tfs = np.concatenate([b[1] for b in self._batches])
            counts = np.concatenate([b[2] for b in self._batches])
            dl = np.concatenate([b[3] for b in self._batches]).astype(np.float32)
        else:
            keys, tfs = np.empty(0, np.uint64), np.empty(0, np.int32)
            counts, dl = np.empty(0, np.int64), np.empty(0, np.float32)
        doc_of = np.repeat(np.arange(n, dtype=np.int64), counts)
        keep = live[doc_of]
        doc_of, keys, tfs = doc_of[keep], keys[keep], tfs[keep]
        term_keys, terms = np.unique(keys, return_inverse=True)
        order = np.lexsort((doc_of, terms))
        terms, doc_of, tfs = terms[order], doc_of[order], tfs[order]
        df = np.bincount(terms, minlength=term_keys.size).astype(np.int64)
        post_ptr = np.concatenate([[0], np.cumsum(df)]).astype(np.int64)
        n_live = int(live.sum())
        self._sum_dl = float(dl[live].sum()) if n_live else 0.0
        avgdl = float(dl[live].mean()) if n_live and dl[live].sum() > 0 else 1.0
        return term_keys, df, post_ptr, doc_of.astype(np.int32), tfs.astype(np.int32), dl, n_live, avgdl

    def _rebuild(self) -> None:
        import torch
        self._term_keys, self._df, self._post_ptr, doc_ids, tfs, dl, self._n_live, self._avgdl = self._host_csr()
        dev = torch.device("cuda", self.device)
        self._doc_ids_dev = torch.from_numpy(doc_ids).to(dev)
        self._tfs_dev = torch.from_numpy(tfs).to(dev)
        self._doc_len_dev = torch.from_numpy(dl).to(dev)
        self._dirty = False

    # ---- scoring -------------------------------------------------------------------------------
    def _query_keys(self, query) -> np.ndarray:
        """Term keys of the query, one per token occurrence, in query order."""
        if isinstance(query, str):
            if self._fast:
                lib = _text_lib()
                toks = [t.encode("ascii") for t in default_tokenize(query)]
                return np.asarray([lib.archi_text_term_key(t, len(t)) for t in toks], dtype=np.uint64)
            known = [self._vocab[t] for t in self.tokenize(query) if t in self._vocab]
            return np.asarray(known, dtype=np.uint64)
        return np.asarray(query).reshape(-1).astype(np.int64).astype(np.uint64)

    def query_terms(self, query) -> List[int]:
        """Term ids (positions in the sorted key array of the last rebuild) of the query's tokens that
        occur in the index; repeated tokens are kept."""
        keys = self._query_keys(query)
        if keys.size == 0 or self._term_keys.size == 0:
            return []
        idx = np.searchsorted(self._term_keys, keys)
        idx[idx >= self._term_keys.size] = 0
        hit = self._term_keys[idx] == keys
        return [int(t) for t in idx[hit]]

    def idf(self, term: int) -> float:
        """idf of a term of this index over the whole table (every index sharing ``self.table``)."""
        if not (0 <= term < self._df.size):
            df = 0
        elif len(self.table.members) == 1:
            df = int(self._df[term])
        else:
            df = int(self.table.df_of_keys(self._term_keys[term:term + 1])[0])
        n = self.table.n_live() if len(self.table.members) > 1 else self._n_live
        return math.log(1.0 + (n - df + 0.5) / (df + 0.5))

    def avgdl(self) -> float:
        return self.table.avgdl() if len(self.table.members) > 1 else self._avgdl

    def posting_ranges(self, query):
        """(starts int64 [t], ends int64 [t], idf float32 [t]) of the query's term occurrences that have postings
        in this index -- what archi_hybrid_search_terms / archi_bm25_accumulate take.  idf uses the table-wide
        statistics."""
        self.table.refresh()
        if self._dirty:
            self._rebuild()
        terms = np.asarray([t for t in self.query_terms(query) if 0 <= t < self._df.size and self._df[t] > 0], dtype=np.int64)
        if terms.size == 0:
            z = np.empty(0, dtype=np.int64)
            return z, z.copy(), np.empty(0, dtype=np.float32)
        if len(self.table.members) == 1:
            df, n = self._df[terms].astype(np.float64), float(self._n_live)
        else:
            df, n = self.table.df_of_keys(self._term_keys[terms]).astype(np.float64), float(self.table.n_live())
        idf = np.log(1.0 + (n - df + 0.5) / (df + 0.5)).astype(np.float32)
        return self._post_ptr[terms].astype(np.int64), self._post_ptr[terms + 1].astype(np.int64), idf

    def device_arrays(self):
        """(doc_ids int32, tfs int32, doc_len float32) CUDA tensors of the posting lists (built on demand)."""
        self.table.refresh()
        if self._dirty:
            self._rebuild()
        return self._doc_ids_dev, self._tfs_dev, self._doc_len_dev

    def score(self, query, out=None):
        """Dense fp32 [rows] BM25 scores of ``query`` on the GPU, 0 where no term matches."""
        import torch
        self.table.refresh()
        if self._dirty:
            self._rebuild()
        n = self._n_docs
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
            self.avgdl(), self.k1, self.b, self.sign, ctypes.c_void_p(out.data_ptr()), stream))
        return out
