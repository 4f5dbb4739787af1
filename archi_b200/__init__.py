"""archi_b200 -- B200-native (sm_100a) retrieval hot path behind archi's vectorstore surface.

    from archi_b200 import B200VectorStore, B200Embeddings, HybridRetriever

Everything numeric runs in libarchi_b200.so (hand-written CUDA, see include/archi_b200.h); there is
no CPU fallback.  Build with ``python -m archi_b200.build``.
"""
from .vectorstore import B200VectorStore, Document
from .retrievers import HybridRetriever, SemanticRetriever, GradingRetriever
from .store import NativeStore, pool_normalize, merge_topk
from .bm25 import LexicalIndex
from .sharded import ShardedStore, PeerExchange, plan_row_shards
from .ingest import IngestionDriver, IngestReport, split_text


def __getattr__(name):
    if name == "B200Embeddings":  # imports torch + transformers: keep it lazy
        from .embeddings import B200Embeddings
        return B200Embeddings
    raise AttributeError(name)


__all__ = ["B200VectorStore", "Document", "HybridRetriever", "SemanticRetriever", "GradingRetriever",
           "NativeStore", "pool_normalize", "merge_topk", "LexicalIndex", "ShardedStore", "PeerExchange",
           "plan_row_shards", "IngestionDriver", "IngestReport", "split_text", "B200Embeddings"]
