"""Shard exchange in isolation: the one-kernel peer-memory exchange (archi_exchange_merge_topk) against NCCL all-gather +
merge kernel, for the record of a (nq, k) search, pipelined (no host sync in the loop) and step by step.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/time_exchange.py [--nq 1024 --k 10 --iters 200]

Optionally a GEMM of about --work-us microseconds precedes every exchange (a stand-in for the local search).
Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from archi_b200.sharded import PeerExchange
from archi_b200.store import merge_topk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--work-us", type=float, default=0.0)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    nq, k = a.nq, a.k
    n = nq * k
    rec_bytes = (n * 12 + 15) // 16 * 16
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    ids = (torch.randperm(n, generator=g).view(nq, k) + rank * n).to(torch.int64)
    scores = torch.rand((nq, k), generator=g).sort(dim=1, descending=True).values
    rec = torch.zeros(rec_bytes, dtype=torch.uint8, device=dev)
    rec[:n * 8].view(torch.int64).copy_(ids.view(-1))
    rec[n * 8:n * 12].view(torch.float32).copy_(scores.view(-1))
    ex = PeerExchange(local, rank, world, None, max(rec_bytes, 1 << 16))
    # a GEMM of roughly work_us microseconds
    work = None
    if a.work_us > 0:
        m = 2048
        x = torch.randn((m, m), device=dev, dtype=torch.bfloat16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            x @ x
        e1.record()
        torch.cuda.synchronize()
        per = e0.elapsed_time(e1) / 20 * 1e3
        reps = max(1, int(round(a.work_us / per)))

        def work():
            for _ in range(reps):
                x @ x

    def peer_step():
        return ex.merge_topk(rec, nq, k, True)

    gathered = torch.empty((world, rec_bytes), dtype=torch.uint8, device=dev)

    def nccl_step():
        dist.all_gather_into_tensor(gathered, rec)
        g_ids = gathered[:, :n * 8].view(torch.int64).view(world, nq, k)
        g_sc = gathered[:, n * 8:n * 12].view(torch.float32).view(world, nq, k)
        return merge_topk(g_sc, g_ids, True)

    def timed(step, sync_each):
        for _ in range(10):
            if work:
                work()
            out = step()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        for _ in range(3):
            if work:
                work()
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            if work:
                work()
            out = step()
            if sync_each:
                torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3, out

    res = {"world": world, "nq": nq, "k": k, "record_bytes": rec_bytes, "work_us_requested": a.work_us}
    t_peer, out_p = timed(peer_step, False)
    t_nccl, out_n = timed(nccl_step, False)
    same = bool(torch.equal(out_p[0], out_n[0]) and torch.equal(out_p[1], out_n[1]))
    res.update(peer_pipelined_us=t_peer, nccl_pipelined_us=t_nccl, same_result=same)
    t_peer_s, _ = timed(peer_step, True)
    t_nccl_s, _ = timed(nccl_step, True)
    res.update(peer_sync_each_us=t_peer_s, nccl_sync_each_us=t_nccl_s)
    if work:
        t_work, _ = timed(lambda: None, False)
        res["work_alone_us"] = t_work
    if rank == 0:
        print(json.dumps(res), flush=True)
    ex.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
