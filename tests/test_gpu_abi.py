"""GPU tests of the C-ABI contract itself: error codes and messages, thread safety of a shared handle,
zero-norm rows, bf16 snapshots."""
import ctypes
import threading

import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_search import check_against_truth, make_store, unit_rows

pytestmark = pytest.mark.gpu


def test_error_codes_and_messages():
    from archi_b200 import _native as N
    from archi_b200.store import NativeStore
    L = N.lib()
    h = ctypes.c_void_p()
    assert L.archi_store_create(0, 384, 7, N.F32, 0, ctypes.byref(h)) == -1           # ARCHI_EINVAL
    assert b"distance_metric must be one of" in L.archi_last_error()
    assert L.archi_store_create(0, 0, N.COSINE, N.F32, 0, ctypes.byref(h)) == -1
    assert L.archi_store_create(99, 8, N.COSINE, N.F32, 0, ctypes.byref(h)) == -1    # no such device
    with pytest.raises(ValueError, match="distance_metric must be one of"):
        NativeStore(8, "manhattan")
    s = NativeStore(8)
    s.append(np.eye(8, dtype=np.float32))
    with pytest.raises(ValueError, match="query dimension"):
        s.search(np.zeros((1, 9), dtype=np.float32), 3)
    q = np.zeros((1, 8), dtype=np.float32)
    sc, ids = np.empty((1, 3), np.float32), np.empty((1, 3), np.int64)
    rc = L.archi_search(s._h, q.ctypes.data_as(ctypes.c_void_p), N.HOST, 1, -1, None, 0, 0,
                        sc.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p), N.HOST, 0, None)
    assert rc == -1
    with pytest.raises(N.NativeError) as e:                                            # tensor path: k <= 128 only
        s.search(np.ones((4, 8), dtype=np.float32), 200, path=N.PATH_TENSOR)
    assert e.value.code == N.EUNSUPPORTED
    with pytest.raises(N.NativeError, match="outside"):
        s.read_rows(5, 10)
    # nq = 0 and k = 0 are no-ops
    assert s.search(np.zeros((0, 8), dtype=np.float32), 3)[1].shape == (0, 3)
    s.close()


def test_shared_handle_is_thread_safe():
    rng = np.random.default_rng(12)
    corpus = unit_rows(rng, 30000, 64)
    s = make_store(corpus)
    queries = [unit_rows(rng, n, 64) for n in (1, 3, 40, 1, 130, 2, 1, 64)]
    want = [s.search(q, 7) for q in queries]
    got = [None] * len(queries)

    def work(i):
        for _ in range(5):
            got[i] = s.search(queries[i], 7)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(queries))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for w, g in zip(want, got):
        assert np.array_equal(w[1], g[1]) and np.allclose(w[0], g[0], rtol=1e-6, atol=1e-7)
    s.close()


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_zero_norm_rows_never_match_under_cosine(storage):
    rng = np.random.default_rng(2)
    corpus = unit_rows(rng, 5000, 32)
    corpus[::7] = 0.0                                     # pgvector yields NaN for these (sorted last)
    s = make_store(corpus, "cosine", storage)
    for q in (unit_rows(rng, 1, 32), unit_rows(rng, 33, 32)):   # streaming and tensor paths
        scores, ids = s.search(q, 20)
        assert (ids >= 0).all() and (ids % 7 != 0).all()
        check_against_truth("cosine", orc.bf16_bits_to_f32(orc.f32_to_bf16_bits(corpus)) if storage == "bf16" else corpus,
                            q, 20, scores, ids, 2e-3 if storage == "bf16" else 1e-5, mask=np.arange(5000) % 7 != 0)
    s.close()


def test_bf16_snapshot_roundtrip(tmp_path):
    from archi_b200.store import NativeStore
    rng = np.random.default_rng(9)
    corpus = unit_rows(rng, 4000, 72)
    s = make_store(corpus, "inner_product", "bf16")
    q = unit_rows(rng, 20, 72)
    before = s.search(q, 5)
    s.save(str(tmp_path / "s.bin"))
    s.close()
    t = NativeStore.load(str(tmp_path / "s.bin"))
    assert (t.storage_dtype, t.metric, t.rows()) == ("bf16", "inner_product", 4000)
    after = t.search(q, 5)
    assert np.array_equal(before[1], after[1]) and np.array_equal(before[0], after[0])
    t.close()


def test_exchange_abi_single_rank():
    """archi_exchange_* with world = 1: the kernel pushes the record into its own buffer, signals
    itself and 'merges' one list -- the output must be the input list (repeated calls alternate the two
    buffer parities); argument errors are reported, not executed."""
    import torch
    from archi_b200 import _native as N
    L = N.lib()
    nq, k = 37, 10
    n = nq * k
    rec = (n * 12 + 15) // 16 * 16
    h = ctypes.c_void_p()
    assert L.archi_exchange_create(0, 0, 1, rec, ctypes.byref(h)) == 0, L.archi_last_error()
    assert L.archi_exchange_create(0, 3, 2, rec, ctypes.byref(ctypes.c_void_p())) == -1       # rank >= world
    raw = (ctypes.c_ubyte * 64)()
    assert L.archi_exchange_local_handle(h, raw) == 0
    rng = np.random.default_rng(8)
    stream = ctypes.c_void_p(int(torch.cuda.current_stream().cuda_stream))
    for call in range(4):
        for larger in (1, 0):
            sc = rng.standard_normal((nq, k)).astype(np.float32)
            sc = -np.sort(-sc, axis=1) if larger else np.sort(sc, axis=1)
            ids = rng.permutation(n).reshape(nq, k).astype(np.int64)
            sc[5, 6:], ids[5, 6:] = np.nan, -1                          # a short list
            mine = torch.zeros(rec, dtype=torch.uint8, device="cuda")
            mine[:n * 8].view(torch.int64).view(nq, k).copy_(torch.from_numpy(ids))
            mine[n * 8:n * 12].view(torch.float32).view(nq, k).copy_(torch.from_numpy(sc))
            out_s = torch.empty((nq, k), dtype=torch.float32, device="cuda")
            out_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
            rc = L.archi_exchange_merge_topk(h, ctypes.c_void_p(mine.data_ptr()), nq, k, larger,
                                             ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()), stream)
            assert rc == 0, L.archi_last_error()
            timed_out = ctypes.c_int(-1)
            assert L.archi_exchange_status(h, ctypes.byref(timed_out)) == 0 and timed_out.value == 0
            assert np.array_equal(out_i.cpu().numpy(), ids)
            assert np.array_equal(out_s.cpu().numpy(), sc, equal_nan=True)
    too_big = torch.zeros(4 * rec, dtype=torch.uint8, device="cuda")
    rc = L.archi_exchange_merge_topk(h, ctypes.c_void_p(too_big.data_ptr()), nq * 4, k, 1,
                                     ctypes.c_void_p(too_big.data_ptr()), ctypes.c_void_p(too_big.data_ptr()), stream)
    assert rc == -1 and b"exceeds the slot" in L.archi_last_error()
    assert L.archi_exchange_destroy(h) == 0
