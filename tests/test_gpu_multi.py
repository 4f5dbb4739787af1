"""Multi-GPU parity: corpus row-sharded over 2 GPUs (one process per GPU, NCCL all-gather of the
per-shard k-lists, device merge) against the oracle on the whole corpus.  Skipped on a 1-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from archi_b200.sharded import ShardedStore, plan_row_shards
    from archi_b200.store import NativeStore
    from oracle import oracle as orc
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    ok = True
    for metric, storage, nq, k in (("cosine", "f32", 5, 10), ("l2", "bf16", 40, 7), ("inner_product", "f32", 130, 100)):
        rng = np.random.default_rng(7)
        corpus = rng.standard_normal((30001, 96)).astype(np.float32)
        queries = rng.standard_normal((nq, 96)).astype(np.float32)
        first, cnt = plan_row_shards(corpus.shape[0], world)[rank]
        store = NativeStore(96, metric, storage, device=rank)
        store.append(corpus[first:first + cnt])
        sh = ShardedStore(store)
        sh.sync_layout(device=torch.device("cuda", rank))
        ok = ok and sh.id_offset == first and sh.total_rows == corpus.shape[0]
        s, i = sh.search(torch.from_numpy(queries).cuda(), k)
        torch.cuda.synchronize()
        s, i = s.cpu().numpy(), i.cpu().numpy()
        stored = orc.bf16_bits_to_f32(orc.f32_to_bf16_bits(corpus)) if storage == "bf16" else corpus
        d_true, i_true = orc.exact_topk(metric, stored, queries, k)
        rel = 2e-3 if storage == "bf16" else 1e-5
        for q in range(nq):
            ok = ok and orc.same_topk_up_to_ties(i[q].tolist(), i_true[q], d_true[q], rel_tol=2e-6, abs_tol=1e-7)
        ok = ok and np.allclose(s, orc.score_from_distance(metric, d_true), rtol=rel, atol=1e-5)
        store.close()
    open(os.path.join(tmp, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    dist.destroy_process_group()


def test_sharded_search_nccl_world2(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]
