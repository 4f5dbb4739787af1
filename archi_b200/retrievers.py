"""Retriever mirrors: the k / score semantics the pipelines rely on.

Reference: src/data_manager/vectorstore/retrievers/{hybrid,semantic,grading}_retriever.py and
utils.py.  The reference classes are LangChain ``BaseRetriever`` pydantic models; when langchain
is importable the reference's own classes work unchanged on a B200VectorStore (it subclasses
``VectorStore``).  These mirrors carry the same defaults and policies without the dependency.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Tuple

from .vectorstore import Document

logger = logging.getLogger(__name__)

# retrievers/utils.py:7-20
INSTRUCTION_AWARE_MODELS = ["Qwen/Qwen3-Embedding-0.6B", "Qwen/Qwen3-Embedding-4B", "Qwen/Qwen3-Embedding-8B"]


def supports_instructions(embedding_name: str, dm_config: Dict[str, Any]) -> Tuple[str, bool]:
    embedding_kwargs = dm_config["embedding_class_map"][embedding_name]["kwargs"]
    embedding_model = embedding_kwargs.get("model") or embedding_kwargs.get("model_name")
    return embedding_model, embedding_model in INSTRUCTION_AWARE_MODELS


def make_instruction_query(instructions: str, query: str) -> str:
    return f"Instruct: {instructions}\nQuery:{query}"


class _Retriever:
    def invoke(self, query: str, **_: Any):
        return self._get_relevant_documents(query)

    get_relevant_documents = invoke


class HybridRetriever(_Retriever):
    """hybrid_retriever.py:20-105: k=5, weights 0.5/0.5 by default; delegates to
    ``vectorstore.hybrid_search``; a RuntimeError whose message says the backend does not support
    hybrid search falls back to semantic-only, any other RuntimeError is re-raised."""

    def __init__(self, vectorstore: Any, k: int = 5, bm25_weight: float = 0.5, semantic_weight: float = 0.5, **kwargs: Any):
        self.vectorstore, self.k, self.bm25_weight, self.semantic_weight = vectorstore, k, bm25_weight, semantic_weight
        self._has_hybrid = hasattr(vectorstore, "hybrid_search")

    def _get_relevant_documents(self, query: str, *, run_manager: Any = None) -> List[Tuple[Document, float]]:
        if self._has_hybrid:
            try:
                return self.vectorstore.hybrid_search(query=query, k=self.k, semantic_weight=self.semantic_weight,
                                                      bm25_weight=self.bm25_weight)
            except RuntimeError as exc:
                message = str(exc).lower()
                if "not supported" in message or "unsupported" in message or "not implemented" in message:
                    logger.warning("Hybrid search not supported by backend, falling back to semantic-only: %s", exc)
                else:
                    raise
        return self.vectorstore.similarity_search_with_score(query, k=self.k)


class SemanticRetriever(_Retriever):
    """semantic_retriever.py:12-46: k=3; optional Qwen3 "Instruct:" prefix; returns (doc, score)."""

    def __init__(self, vectorstore: Any, dm_config: Dict[str, Any], k: int = 3, instructions: Optional[str] = None):
        self.vectorstore, self.dm_config, self.k, self.instructions = vectorstore, dm_config, k, instructions

    def _get_relevant_documents(self, query: str) -> List[Tuple[Document, float]]:
        embedding_name = self.dm_config["embedding_name"]
        embedding_model, supported = supports_instructions(embedding_name, self.dm_config)
        if self.instructions and supported:
            query = make_instruction_query(self.instructions, query)
        elif self.instructions:
            logger.warning("Instructions provided but model '%s' not in supported models: %s", embedding_model,
                           INSTRUCTION_AWARE_MODELS)
        return self.vectorstore.similarity_search_with_score(query, k=self.k)


class GradingRetriever(_Retriever):
    """grading_retriever.py:11-25: k=3; bare documents."""

    def __init__(self, vectorstore: Any, k: int = 3):
        self.vectorstore, self.k = vectorstore, k

    def _get_relevant_documents(self, query: str) -> List[Document]:
        return self.vectorstore.similarity_search(query, k=self.k)
