"""ShardedStore -- one corpus row-sharded over the GPUs of a box, one process per GPU.

The reference has no distributed path (single PostgreSQL backend per query).  Here each rank owns
a contiguous block of rows in its own HBM (a NativeStore), every rank receives the full query
batch, computes its local exact top-k, and the per-shard k-lists are exchanged with ONE NCCL
all-gather (``Q*k*12`` bytes per rank over NVLink) and merged on the device by
``archi_merge_topk``.  Exact because the global top-k is a subset of the union of the local ones.

The exchange step is the only collective; there is none on the data path of the scan itself.

On one box the exchange does not go through NCCL at all when the ranks can map each other's memory
(CUDA IPC): ``PeerExchange`` pushes every rank's record straight into its peers' HBM over NVLink and
merges in the same kernel (csrc/exchange.cu).  It is validated against the NCCL path when it is set up
and switched off (with a warning) if the box cannot do it; ``ARCHI_PEER_EXCHANGE=0`` forces NCCL.
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


class PeerExchange:
    """Handle of archi_exchange_* (include/archi_b200.h): one gather buffer per rank, mapped by every
    peer.  Construction is collective over ``group``; raises RuntimeError on every rank if any rank
    cannot map its peers."""

    def __init__(self, device_index: int, rank: int, world: int, group, max_record_bytes: int):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _native as N
        self._N, self._ctypes, self._dist, self.group = N, ctypes, dist, group
        self.device_index, self.rank, self.world = int(device_index), int(rank), int(world)
        self.max_record_bytes = int(max_record_bytes)
        L = N.lib()
        dev = torch.device("cuda", self.device_index)
        self._h = None

        def everyone_ok(ok: bool) -> bool:
            # every step that can fail locally is followed by a collective verdict, so that no rank is
            # left waiting in a collective the failing rank never enters
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return int(t.item()) == 1

        why = ""
        raw = (ctypes.c_ubyte * 64)()
        h = ctypes.c_void_p()
        if L.archi_exchange_create(self.device_index, self.rank, self.world, self.max_record_bytes, ctypes.byref(h)) == 0:
            self._h = h
            if L.archi_exchange_local_handle(h, raw) != 0:
                why = L.archi_last_error().decode("utf-8", "replace")
        else:
            why = L.archi_last_error().decode("utf-8", "replace")
        if not everyone_ok(why == ""):
            self.close(barrier=False)
            raise RuntimeError("peer-memory exchange unavailable: " + (why or "a peer rank could not export its buffer"))
        mine = torch.tensor(list(bytes(raw)), dtype=torch.uint8, device=dev)
        handles = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(handles, mine, group=group)
        blob = handles.cpu().numpy().tobytes()
        if L.archi_exchange_connect(h, blob) != 0:
            why = L.archi_last_error().decode("utf-8", "replace")
        if not everyone_ok(why == ""):
            self.close(barrier=False)
            raise RuntimeError("peer-memory exchange unavailable: " + (why or "a peer rank could not map this rank's buffer"))

    def merge_topk(self, record, nq: int, k: int, larger_is_better: bool):
        """record: uint8 CUDA tensor {ids [nq,k] int64 | scores [nq,k] fp32} of this rank.  Returns the
        merged (scores [nq,k], ids [nq,k]) of all ranks; stream-ordered, nothing synchronised."""
        import torch
        ctypes, N = self._ctypes, self._N
        dev = record.device
        out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        stream = ctypes.c_void_p(int(torch.cuda.current_stream(dev).cuda_stream))
        N.check(N.lib().archi_exchange_merge_topk(self._h, ctypes.c_void_p(record.data_ptr()), int(nq), int(k),
                                                  int(larger_is_better), ctypes.c_void_p(out_s.data_ptr()),
                                                  ctypes.c_void_p(out_i.data_ptr()), stream))
        return out_s, out_i

    def timed_out(self) -> bool:
        """Synchronises the device; True if any call gave up waiting for a peer."""
        flag = self._ctypes.c_int(0)
        self._N.check(self._N.lib().archi_exchange_status(self._h, self._ctypes.byref(flag)))
        return flag.value != 0

    def close(self, barrier: bool = True) -> None:
        if getattr(self, "_h", None) is None:
            return
        if barrier:
            import torch
            torch.cuda.synchronize(self.device_index)
            self._dist.barrier(group=self.group)      # nobody may still be pushing into a buffer about to be freed
        self._N.lib().archi_exchange_destroy(self._h)
        self._h = None


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
        self._records = {}              # (nq, k, device) -> packed record buffer of _search_packed
        self._exchange = None           # PeerExchange, or False once found unavailable
        self.exchange_kind = "nccl"     # what search() uses between the ranks: "nccl" | "peer-memory"

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
        """CUDA path: the local top-k lands directly in this rank's record {ids | scores | pad to 16 B};
        ONE all-gather moves every record, and the merge kernel reads the gathered buffer in place."""
        import torch
        nq = 1 if queries.dim() == 1 else int(queries.shape[0])
        n = nq * k
        rec = (n * 12 + 15) // 16 * 16      # the exchange kernel moves 16-byte words
        # this rank's record is scratch between the local search and the exchange of one call (stream-ordered):
        # one buffer per shape, not one allocation per search
        key = (nq, k, queries.device)
        mine = self._records.get(key)
        if mine is None:
            if len(self._records) > 16:
                self._records.clear()
            mine = self._records[key] = torch.empty(rec, dtype=torch.uint8, device=queries.device)
        ids = mine[:n * 8].view(torch.int64).view(nq, k)
        scores = mine[n * 8:n * 12].view(torch.float32).view(nq, k)
        self._local_search(queries, k, self.id_offset, out=(scores, ids))
        exchange = self._peer_exchange(queries.device, rec)
        if exchange is not None:
            return exchange.merge_topk(mine, nq, k, self.larger_is_better)
        gathered = torch.empty((self.world, rec), dtype=torch.uint8, device=queries.device)
        self._dist.all_gather_into_tensor(gathered, mine, group=self.group)
        g_ids = gathered[:, :n * 8].view(torch.int64).view(self.world, nq, k)
        g_scores = gathered[:, n * 8:n * 12].view(torch.float32).view(self.world, nq, k)
        return self._merge(g_scores, g_ids, self.larger_is_better)

    # ---- peer-memory exchange (one box, CUDA IPC) --------------------------------------------------
    def _peer_exchange(self, device, rec_bytes: int):
        """The PeerExchange to use for records of ``rec_bytes``, or None (NCCL).  Collective: every rank
        reaches the same decision at the same call."""
        import os
        if self._exchange is False:
            return None
        if self._exchange is not None and self._exchange.max_record_bytes >= rec_bytes:
            return self._exchange
        if os.environ.get("ARCHI_PEER_EXCHANGE", "1") == "0":
            self._exchange = False
            return None
        if self._exchange is not None:
            self._exchange.close()
            self._exchange = None
        try:
            ex = PeerExchange(device.index, self.rank, self.world, self.group, max(rec_bytes, 1 << 16))
            self._validate_exchange(ex, device)
            self._exchange, self.exchange_kind = ex, "peer-memory"
        except RuntimeError as e:
            if self.rank == 0:
                import warnings
                warnings.warn(f"archi_b200: {e}; the shard exchange stays on NCCL")
            self._exchange, self.exchange_kind = False, "nccl"
            return None
        return self._exchange

    def _validate_exchange(self, ex: "PeerExchange", device) -> None:
        """One exchange of synthetic lists, checked against NCCL all-gather + archi_merge_topk_strided."""
        import torch
        nq, k = 5, 7
        n = nq * k
        rec = (n * 12 + 15) // 16 * 16
        g = torch.Generator(device="cpu").manual_seed(1234 + self.rank)
        sc = torch.sort(torch.rand((nq, k), generator=g), dim=1, descending=self.larger_is_better).values
        mine = torch.zeros(rec, dtype=torch.uint8, device=device)
        mine[:n * 8].view(torch.int64).view(nq, k).copy_(torch.arange(n).view(nq, k) + 1000 * self.rank)
        mine[n * 8:n * 12].view(torch.float32).view(nq, k).copy_(sc)
        bad = False
        try:
            got_s, got_i = ex.merge_topk(mine, nq, k, self.larger_is_better)
        except RuntimeError:
            bad = True
        # the collectives below are entered by every rank whatever happened locally
        gathered = torch.empty((self.world, rec), dtype=torch.uint8, device=device)
        self._dist.all_gather_into_tensor(gathered, mine, group=self.group)
        try:
            want_s, want_i = self._merge(gathered[:, n * 8:n * 12].view(torch.float32).view(self.world, nq, k),
                                         gathered[:, :n * 8].view(torch.int64).view(self.world, nq, k), self.larger_is_better)
            bad = bad or ex.timed_out() or not (torch.equal(got_i, want_i) and torch.equal(got_s, want_s))
        except RuntimeError:
            bad = True
        flag = torch.tensor([1 if bad else 0], dtype=torch.int32, device=device)
        self._dist.all_reduce(flag, op=self._dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()):
            ex.close()
            raise RuntimeError("peer-memory exchange failed its self-check against the NCCL path")

    def check(self) -> None:
        """Synchronises; raises if a peer-memory exchange ever timed out (its results were invalid)."""
        if self._exchange and self._exchange.timed_out():
            raise RuntimeError("archi_b200: a peer-memory shard exchange timed out waiting for a rank")

    def close(self) -> None:
        if self._exchange:
            self._exchange.close()
        self._exchange = None

