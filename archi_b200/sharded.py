"""ShardedStore -- one corpus row-sharded over the GPUs of a box, one process per GPU.

The reference has no distributed path (single PostgreSQL backend per query).  Here each rank owns
a contiguous block of rows in its own HBM (a NativeStore), every rank receives the full query
batch, computes its local exact top-k, and the per-shard k-lists are exchanged with ONE NCCL
all-gather (``Q*k*12`` bytes per rank over NVLink) and merged on the device by
``archi_merge_topk``.  Exact because the global top-k is a subset of the union of the local ones.

The exchange step is the only collective; there is none on the data path of the scan itself.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def plan_row_shards(n_rows: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous row blocks [(first, count)] per rank: rows_g = ceil(n / G), the last ranks may be
    short or empty.  Global id = first + local id."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    per = -(-n_rows // world_size) if n_rows > 0 else 0
    out = []
    for r in range(world_size):
        first = min(r * per, n_rows)
        out.append((first, max(0, min(per, n_rows - first))))
    return out


def offsets_from_counts(counts: Sequence[int]) -> List[int]:
    """Exclusive prefix sum: the id offset of each rank's block."""
    out, acc = [], 0
    for c in counts:
        out.append(acc)
        acc += int(c)
    return out


class ShardedStore:
    """``local_search(queries, k, id_offset) -> (scores [nq,k], ids [nq,k])`` and
    ``merge(scores [G,nq,k], ids [G,nq,k], larger_is_better) -> (scores, ids)`` default to the
    CUDA implementations; tests on CPU (gloo) inject stand-ins to exercise the host logic."""

    def __init__(self, native=None, *, group=None, larger_is_better: Optional[bool] = None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 local_rows: Optional[Callable[[], int]] = None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.native = native
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if larger_is_better is None:
            larger_is_better = native.larger_is_better
        self.larger_is_better = bool(larger_is_better)
        self._local_search = local_search or self._native_search
        self._merge = merge or self._native_merge
        self._local_rows = local_rows or (lambda: self.native.rows())
        self.id_offset = 0
        self.total_rows = 0

    # ---- defaults: the CUDA path -------------------------------------------------------------------
    def _native_search(self, queries, k, id_offset, out=None):
        return self.native.search(queries, k, id_offset=id_offset, out=out)

    @staticmethod
    def _native_merge(scores, ids, larger_is_better):
        from .store import merge_topk
        return merge_topk(scores, ids, larger_is_better)

    # ---- layout ----------------------------------------------------------------------------------------
    def sync_layout(self, device=None) -> None:
        """All-gather the local row counts and derive this rank's global id offset."""
        import torch
        n = int(self._local_rows())
        if self.world == 1:
            self.id_offset, self.total_rows = 0, n
            return
        t = torch.tensor([n], dtype=torch.int64, device=device)
        gathered = [torch.zeros_like(t) for _ in range(self.world)]
        self._dist.all_gather(gathered, t, group=self.group)
        counts = [int(g.item()) for g in gathered]
        self.id_offset = offsets_from_counts(counts)[self.rank]
        self.total_rows = sum(counts)

    # ---- search ------------------------------------------------------------------------------------------
    def search(self, queries, k: int):
        """Every rank passes the same ``queries`` and receives the same merged (scores, ids)."""
        import torch
        if self.world > 1 and self._local_search == self._native_search and self._merge == self._native_merge \
                and getattr(queries, "is_cuda", False):
            return self._search_packed(queries, int(k))
        scores, ids = self._local_search(queries, k, self.id_offset)
        if self.world == 1:
            return scores, ids
        g_scores = torch.empty((self.world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
        g_ids = torch.empty((self.world,) + tuple(ids.shape), dtype=ids.dtype, device=ids.device)
        if scores.is_cuda:
            self._dist.all_gather_into_tensor(g_scores, scores.contiguous(), group=self.group)
            self._dist.all_gather_into_tensor(g_ids, ids.contiguous(), group=self.group)
        else:  # gloo
            ls = [g_scores[r] for r in range(self.world)]
            li = [g_ids[r] for r in range(self.world)]
            self._dist.all_gather(ls, scores.contiguous(), group=self.group)
            self._dist.all_gather(li, ids.contiguous(), group=self.group)
        return self._merge(g_scores, g_ids, self.larger_is_better)

    def _search_packed(self, queries, k: int):
        """CUDA path: the local top-k lands directly in this rank's record {ids | scores | pad to 8 B};
        ONE all-gather moves every record, and the merge kernel reads the gathered buffer in place."""
        import torch
        nq = 1 if queries.dim() == 1 else int(queries.shape[0])
        n = nq * k
        rec = (n * 12 + 7) // 8 * 8
        mine = torch.empty(rec, dtype=torch.uint8, device=queries.device)
        ids = mine[:n * 8].view(torch.int64).view(nq, k)
        scores = mine[n * 8:n * 12].view(torch.float32).view(nq, k)
        self._local_search(queries, k, self.id_offset, out=(scores, ids))
        gathered = torch.empty((self.world, rec), dtype=torch.uint8, device=queries.device)
        self._dist.all_gather_into_tensor(gathered, mine, group=self.group)
        g_ids = gathered[:, :n * 8].view(torch.int64).view(self.world, nq, k)
        g_scores = gathered[:, n * 8:n * 12].view(torch.float32).view(self.world, nq, k)
        return self._merge(g_scores, g_ids, self.larger_is_better)
