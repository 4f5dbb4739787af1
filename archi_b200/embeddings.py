"""B200Embeddings -- the embedding function behind add_documents / similarity_search.

Mirrors what the reference resolves for ``embedding_name: HuggingFaceEmbeddings``
(src/utils/config_service.py:479-485; base-config.yaml:143-152): langchain-huggingface's
HuggingFaceEmbeddings over sentence-transformers' Transformer -> Pooling(mean) -> Normalize
[external].  Here the encoder forward stays in PyTorch (on the GPU) and the Pooling + Normalize
tail is the fused sm_100a kernel ``archi_pool_normalize`` (archi_b200/csrc/pool.cu); for
``add_documents`` the same kernel writes the rows straight into the corpus matrix.

No model weights or vocabularies can be downloaded in the build environment: when the named model
is not in the local HF cache the encoder is a random-init ``BertModel`` of the MiniLM-L6 shape
(6 layers x 384 hidden x 12 heads, 22.7 M parameters) and the tokenizer a hashing word tokenizer.
Both can be injected (``model=``, ``tokenizer=``).
"""
from __future__ import annotations

import logging
import zlib
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import store as _store

logger = logging.getLogger(__name__)

MINILM_L6 = dict(vocab_size=30522, hidden_size=384, num_hidden_layers=6, num_attention_heads=12,
                 intermediate_size=1536, max_position_embeddings=512)


class HashTokenizer:
    """Offline stand-in for a WordPiece tokenizer: lower-cased whitespace words hashed into the
    vocabulary, [CLS] ... [SEP], truncated to max_len.  Deterministic across processes."""

    def __init__(self, vocab_size: int = 30522, cls_id: int = 101, sep_id: int = 102, pad_id: int = 0):
        self.vocab_size, self.cls_id, self.sep_id, self.pad_id = vocab_size, cls_id, sep_id, pad_id

    def __call__(self, texts: Sequence[str], max_len: int) -> Tuple[np.ndarray, np.ndarray]:
        rows = []
        for t in texts:
            ids = [1000 + zlib.crc32(w.encode("utf-8")) % (self.vocab_size - 1000) for w in t.lower().split()]
            rows.append([self.cls_id] + ids[: max_len - 2] + [self.sep_id])
        L = max(len(r) for r in rows)
        ids = np.full((len(rows), L), self.pad_id, dtype=np.int64)
        mask = np.zeros((len(rows), L), dtype=np.int64)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = r
            mask[i, : len(r)] = 1
        return ids, mask


class B200Embeddings:
    """``batch_size=None`` (default) batches by TOKEN BUDGET: consecutive texts share an encoder forward until
    ``sequences x padded length`` would exceed ``max_batch_tokens``, and the padded length is rounded up to a multiple
    of 32 -- a forward costs ~20 ms of host time in PyTorch whatever its size, and every new (batch, length) shape
    costs allocator and GEMM-heuristic work, so few large forwards of a handful of shapes is what makes ingestion
    GPU-bound (sentence-transformers' fixed 32 texts per forward is not).  An explicit ``batch_size`` restores
    fixed-count batches padded to the longest member."""

    def __init__(self, model_name: str = "sentence-transformers/all-MiniLM-L6-v2", *, model=None,
                 tokenizer: Optional[Callable] = None, device: int = 0, max_seq_length: int = 256,
                 batch_size: Optional[int] = None, max_batch_tokens: int = 262144, dtype: str = "bf16", seed: int = 0):
        import torch
        self.model_name, self.device, self.max_seq_length, self.batch_size = model_name, int(device), max_seq_length, batch_size
        self.max_batch_tokens = int(max_batch_tokens)
        self._torch_dtype = torch.bfloat16 if dtype == "bf16" else torch.float32
        self._dev = torch.device("cuda", self.device)
        if model is None:
            model = self._load_or_init(model_name, seed)
        self.model = model.to(self._dev, self._torch_dtype).eval()
        self.tokenizer = tokenizer or self._load_tokenizer(model_name)
        self.hidden_size = int(getattr(self.model.config, "hidden_size", 0)) or None

    @staticmethod
    def _load_or_init(model_name: str, seed: int):
        import torch
        from transformers import AutoModel, BertConfig, BertModel
        try:
            return AutoModel.from_pretrained(model_name, local_files_only=True)
        except Exception:
            logger.warning("no local weights for %s: using a random-init MiniLM-L6-shaped BertModel", model_name)
            torch.manual_seed(seed)
            return BertModel(BertConfig(**MINILM_L6), add_pooling_layer=False)

    def _load_tokenizer(self, model_name: str):
        try:
            from transformers import AutoTokenizer
            tok = AutoTokenizer.from_pretrained(model_name, local_files_only=True)

            def call(texts, max_len):
                enc = tok(list(texts), padding=True, truncation=True, max_length=max_len, return_tensors="np")
                return enc["input_ids"].astype(np.int64), enc["attention_mask"].astype(np.int64)
            return call
        except Exception:
            vocab = int(getattr(self.model.config, "vocab_size", 30522))
            return HashTokenizer(vocab_size=vocab)

    # ---- encoder forward (PyTorch) -> last_hidden_state, attention_mask on the GPU ----------------------
    def _forward_ids(self, ids: np.ndarray, mask: np.ndarray):
        import torch
        ids_t = torch.from_numpy(np.ascontiguousarray(ids)).to(self._dev, non_blocking=True)
        mask_t = torch.from_numpy(np.ascontiguousarray(mask)).to(self._dev, non_blocking=True)
        with torch.inference_mode():
            hidden = self.model(input_ids=ids_t, attention_mask=mask_t).last_hidden_state
        return hidden, mask_t

    def _forward(self, texts: Sequence[str]):
        ids, mask = self.tokenizer(texts, self.max_seq_length)
        return self._forward_ids(ids, mask)

    def _batches(self, texts: Sequence[str]):
        """Yields (hidden [b, L, H], mask [b, L]) over consecutive slices of ``texts``, in order."""
        if self.batch_size:
            for s in range(0, len(texts), self.batch_size):
                yield self._forward(texts[s: s + self.batch_size])
            return
        # token budget: tokenise in slices (bounded host memory), then cut each slice where b x padded length
        # would pass the budget; lengths are rounded up to a multiple of 32 (few distinct shapes)
        pad_id = int(getattr(self.tokenizer, "pad_id", 0) or 0)
        for s0 in range(0, len(texts), 4096):
            ids, mask = self.tokenizer(texts[s0: s0 + 4096], self.max_seq_length)
            lens = mask.sum(axis=1)
            n, at = ids.shape[0], 0
            while at < n:
                end, longest = at, 1
                while end < n:
                    cand = max(longest, int(lens[end]))
                    padded = min(-(-cand // 32) * 32, max(self.max_seq_length, cand))
                    if end > at and (end - at + 1) * padded > self.max_batch_tokens:
                        break
                    longest = cand
                    end += 1
                L = min(-(-longest // 32) * 32, max(self.max_seq_length, longest))
                b_ids = np.full((end - at, L), pad_id, dtype=np.int64)
                b_mask = np.zeros((end - at, L), dtype=np.int64)
                w = min(L, ids.shape[1])
                b_ids[:, :w] = ids[at:end, :w]
                b_mask[:, :w] = mask[at:end, :w]
                yield self._forward_ids(b_ids, b_mask)
                at = end

    @staticmethod
    def _clean(texts: Sequence[str]) -> List[str]:
        # langchain-huggingface's HuggingFaceEmbeddings replaces newlines before encoding [external]
        return [t.replace("\n", " ") for t in texts]

    def embed_documents_device(self, texts: Sequence[str]):
        """[n, H] fp32 CUDA tensor of unit-norm embeddings, in input order."""
        import torch
        texts = self._clean(texts)
        outs = []
        for hidden, mask in self._batches(texts):
            out_f32, _ = _store.pool_normalize(hidden, mask)
            outs.append(out_f32)
        return torch.cat(outs, dim=0) if outs else torch.empty((0, self.hidden_size or 0), device=self._dev)

    def embed_documents_into(self, texts: Sequence[str], collection) -> int:
        """Encoder forward -> one fused kernel per batch that pools, normalises, casts and appends
        the rows to ``collection``'s GPU store.  Returns the first row id."""
        texts = self._clean(texts)
        first = None
        for hidden, mask in self._batches(texts):
            native = collection.ensure_native(hidden.shape[-1])
            r0 = native.pool_normalize_append(hidden, mask)
            first = r0 if first is None else first
        return int(first)

    # ---- LangChain Embeddings surface ---------------------------------------------------------------------
    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        return self.embed_documents_device(texts).cpu().tolist()

    def embed_query(self, text: str) -> List[float]:
        return self.embed_documents_device([text])[0].cpu().tolist()

    def embed_query_device(self, text: str):
        """[1, H] fp32 CUDA tensor: the query embedding straight from the pool+normalise kernel.  B200VectorStore
        passes it to the search kernels as a device pointer -- the reference turns the vector into decimal text
        for the SQL statement (postgres_vectorstore.py:313,391)."""
        return self.embed_documents_device([text])[:1]
