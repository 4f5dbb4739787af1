"""GPU parity tests of the search path: CUDA (through the C ABI) against the oracle on the same
seeded inputs.  Bar: top-k id sets identical to the fp64 exact search over the stored values up to
tie order; scores within 1e-5 relative (fp32 storage) / 2e-3 (bf16 storage)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

REL_F32 = 1e-5
REL_BF16 = 2e-3


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def make_store(corpus, metric="cosine", storage="f32"):
    from archi_b200.store import NativeStore
    s = NativeStore(corpus.shape[1], metric, storage)
    assert s.append(corpus) == 0
    return s


def stored_values(corpus, storage):
    return orc.bf16_bits_to_f32(orc.f32_to_bf16_bits(corpus)) if storage == "bf16" else corpus


def check_against_truth(metric, stored, queries, k, scores, ids, rel, mask=None):
    d_true, i_true = orc.exact_topk(metric, stored, queries, k, mask=mask)
    s_true = orc.score_from_distance(metric, d_true)
    kk = d_true.shape[1]
    for q in range(queries.shape[0]):
        got = ids[q, :kk].tolist()
        assert orc.same_topk_up_to_ties(got, i_true[q], d_true[q], rel_tol=2e-6, abs_tol=1e-7), \
            f"query {q}: ids {got} != {i_true[q].tolist()}"
        assert (ids[q, kk:] == -1).all() and np.isnan(scores[q, kk:]).all()
        # scores: compare rank by rank (ties permute ids, not scores)
        assert np.allclose(scores[q, :kk], s_true[q], rtol=rel, atol=rel * 1e-1), \
            f"query {q}: scores {scores[q, :kk]} vs {s_true[q]}"
        # the returned order is best first
        key = scores[q, :kk] if metric == "cosine" else -scores[q, :kk]
        assert (np.diff(key) <= 1e-7).all()


@pytest.mark.parametrize("storage", ["f32", "bf16"])
@pytest.mark.parametrize("metric", ["cosine", "l2", "inner_product"])
@pytest.mark.parametrize("nq", [1, 2, 3, 8, 11])
def test_search_matches_oracle(metric, storage, nq):
    rng = np.random.default_rng(1234 + nq)
    corpus = (unit_rows(rng, 20011, 384) * rng.uniform(0.5, 1.5, size=(20011, 1))).astype(np.float32)
    queries = unit_rows(rng, nq, 384)
    s = make_store(corpus, metric, storage)
    scores, ids = s.search(queries, 10)
    check_against_truth(metric, stored_values(corpus, storage), queries, 10, scores, ids,
                        REL_BF16 if storage == "bf16" else REL_F32)
    s.close()


@pytest.mark.parametrize("dim", [1, 3, 8, 30, 100, 257, 768, 1024, 1536])
def test_dims_including_unaligned(dim):
    rng = np.random.default_rng(dim)
    corpus = rng.standard_normal((3000, dim)).astype(np.float32)
    queries = rng.standard_normal((4, dim)).astype(np.float32)
    for storage in ("f32", "bf16"):
        s = make_store(corpus, "cosine", storage)
        scores, ids = s.search(queries, 5)
        check_against_truth("cosine", stored_values(corpus, storage), queries, 5, scores, ids,
                            REL_BF16 if storage == "bf16" else REL_F32)
        s.close()


@pytest.mark.parametrize("k", [1, 4, 32, 33, 100, 128, 129, 300])
def test_k_values_including_multi_pass(k):
    rng = np.random.default_rng(k)
    corpus = unit_rows(rng, 5000, 64)
    queries = unit_rows(rng, 3, 64)
    s = make_store(corpus)
    scores, ids = s.search(queries, k)
    check_against_truth("cosine", corpus, queries, k, scores, ids, REL_F32)
    for q in range(3):
        assert len(set(ids[q].tolist())) == k
    s.close()


def test_k_larger_than_rows_and_empty_store():
    from archi_b200.store import NativeStore
    rng = np.random.default_rng(0)
    corpus = unit_rows(rng, 7, 16)
    s = make_store(corpus)
    scores, ids = s.search(unit_rows(rng, 2, 16), 10)
    assert (ids[:, :7] >= 0).all() and (ids[:, 7:] == -1).all() and np.isnan(scores[:, 7:]).all()
    for q in range(2):
        assert sorted(ids[q, :7].tolist()) == list(range(7))
    s.close()
    e = NativeStore(16)
    scores, ids = e.search(unit_rows(rng, 1, 16), 4)
    assert (ids == -1).all()
    assert e.count() == 0
    e.close()


def test_ties_prefer_lower_id():
    corpus = np.tile(np.array([[0.6, 0.8, 0.0, 0.0]], dtype=np.float32), (1000, 1))
    corpus[::2] = np.array([0.0, 0.0, 1.0, 0.0], dtype=np.float32)
    s = make_store(corpus)
    scores, ids = s.search(np.array([[0.6, 0.8, 0, 0]], dtype=np.float32), 6)
    assert ids[0].tolist() == [1, 3, 5, 7, 9, 11]
    assert np.allclose(scores[0], 1.0, atol=1e-6)
    s.close()


def test_filter_mask_and_tombstones():
    import torch
    rng = np.random.default_rng(77)
    corpus = unit_rows(rng, 4099, 96)
    queries = unit_rows(rng, 5, 96)
    s = make_store(corpus)
    keep = rng.random(4099) < 0.3
    words = np.concatenate([orc.pack_mask(keep), np.zeros(1, np.uint32)])
    fm = torch.from_numpy(words.view(np.int32).copy()).cuda()
    scores, ids = s.search(queries, 10, filter_mask=fm)
    check_against_truth("cosine", corpus, queries, 10, scores, ids, REL_F32, mask=keep)
    # tombstones: delete the current top-3 of query 0 and search again
    _, top = s.search(queries[:1], 3)
    s.delete_rows(top[0])
    assert s.count() == 4099 - 3 and s.rows() == 4099
    alive = np.ones(4099, dtype=bool)
    alive[top[0]] = False
    scores, ids = s.search(queries, 10)
    check_against_truth("cosine", corpus, queries, 10, scores, ids, REL_F32, mask=alive)
    scores, ids = s.search(queries, 10, filter_mask=fm)
    check_against_truth("cosine", corpus, queries, 10, scores, ids, REL_F32, mask=alive & keep)
    s.delete_rows(top[0])                       # deleting twice changes nothing
    assert s.count() == 4099 - 3
    s.close()


def test_device_tensors_in_and_out_and_id_offset():
    import torch
    rng = np.random.default_rng(5)
    corpus = unit_rows(rng, 9000, 128)
    queries = unit_rows(rng, 6, 128)
    s = make_store(corpus, "inner_product", "bf16")
    sc_h, id_h = s.search(queries, 10)
    sc_d, id_d = s.search(torch.from_numpy(queries).cuda(), 10, id_offset=1_000_000)
    torch.cuda.synchronize()
    assert (id_d.cpu().numpy() == id_h + 1_000_000).all()
    assert np.array_equal(sc_d.cpu().numpy(), sc_h)
    s.close()


def test_append_grows_reads_back_and_snapshot(tmp_path):
    from archi_b200.store import NativeStore
    rng = np.random.default_rng(8)
    a, b = unit_rows(rng, 1500, 48), unit_rows(rng, 700, 48)
    s = NativeStore(48, "cosine", "f32", capacity_rows=16)
    assert s.append(a) == 0 and s.append(b) == 1500
    assert s.rows() == 2200 and s.capacity() >= 2200
    assert np.array_equal(s.read_rows(1400, 300), np.concatenate([a, b])[1400:1700])
    s.delete_rows([5, 1600])
    q = unit_rows(rng, 3, 48)
    before = s.search(q, 9)
    path = str(tmp_path / "shard.bin")
    s.save(path)
    s.close()
    t = NativeStore.load(path)
    assert (t.dim, t.metric, t.storage_dtype, t.rows(), t.count()) == (48, "cosine", "f32", 2200, 2198)
    after = t.search(q, 9)
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    t.reset()
    assert t.rows() == 0 and (t.search(q, 2)[1] == -1).all()
    t.close()


def test_golden_fixtures_through_the_abi(golden_dir):
    g = np.load(os.path.join(golden_dir, "search_2048x96.npz"))
    for metric in orc.METRICS:
        s = make_store(g["corpus"], metric)
        scores, ids = s.search(g["queries"], 10)
        d = g[f"{metric}_dist"]
        for q in range(5):
            assert orc.same_topk_up_to_ties(ids[q].tolist(), g[f"{metric}_ids"][q], d[q], rel_tol=2e-6, abs_tol=1e-7)
        assert np.allclose(scores, orc.score_from_distance(metric, d), rtol=REL_F32, atol=1e-6)
        s.close()
    s = make_store(g["corpus"], "cosine", "bf16")
    scores, ids = s.search(g["queries"], 10)
    for q in range(5):
        assert orc.same_topk_up_to_ties(ids[q].tolist(), g["bf16_cosine_ids"][q], g["bf16_cosine_dist"][q],
                                        rel_tol=2e-6, abs_tol=1e-7)
    s.close()


def test_hybrid_matches_oracle_and_golden(golden_dir):
    import torch
    h = np.load(os.path.join(golden_dir, "hybrid_512x64.npz"))
    bm = torch.from_numpy(np.nan_to_num(h["bm25"], nan=0.0).astype(np.float32)).cuda()[None, :].contiguous()
    for metric in orc.METRICS:
        s = make_store(h["corpus"], metric)
        for ws, wb in ((0.7, 0.3), (0.4, 0.6)):
            scores, ids = s.search(h["query"], 8, bm25=bm, semantic_weight=ws, bm25_weight=wb, hybrid=True)
            assert ids[0].tolist() == h[f"{metric}_{ws}_{wb}_ids"].tolist()
            assert np.allclose(scores[0], h[f"{metric}_{ws}_{wb}_combined"], rtol=1e-5, atol=1e-6)
        s.close()


def test_merge_topk_kernel():
    import torch
    from archi_b200.store import merge_topk
    rng = np.random.default_rng(21)
    G, nq, k = 8, 37, 10
    for larger in (True, False):
        s = rng.standard_normal((G, nq, k)).astype(np.float32)
        s = -np.sort(-s, axis=2) if larger else np.sort(s, axis=2)
        ids = rng.permutation(G * nq * k).reshape(G, nq, k).astype(np.int64)
        s[3, :, 6:] = np.nan                       # a short shard
        ids[3, :, 6:] = -1
        s[5] = np.nan                              # an empty shard
        ids[5] = -1
        ms, mi = merge_topk(torch.from_numpy(s).cuda(), torch.from_numpy(ids).cuda(), larger)
        ms, mi = ms.cpu().numpy(), mi.cpu().numpy()
        flat_s = s.transpose(1, 0, 2).reshape(nq, -1)
        flat_i = ids.transpose(1, 0, 2).reshape(nq, -1)
        key = np.where(flat_i >= 0, flat_s if larger else -flat_s, -np.inf)
        order = np.lexsort((flat_i, -key), axis=1)[:, :k]
        assert np.array_equal(mi, np.take_along_axis(flat_i, order, 1))
        assert np.array_equal(ms, np.take_along_axis(flat_s, order, 1))
        # the same lists as records {ids | scores | pad} of one all-gather buffer (archi_merge_topk_strided)
        n = nq * k
        rec = (n * 12 + 7) // 8 * 8
        packed = torch.zeros((G, rec), dtype=torch.uint8, device="cuda")
        p_ids = packed[:, :n * 8].view(torch.int64).view(G, nq, k)
        p_sc = packed[:, n * 8:n * 12].view(torch.float32).view(G, nq, k)
        p_ids.copy_(torch.from_numpy(ids))
        p_sc.copy_(torch.from_numpy(s))
        assert not p_sc.is_contiguous()
        ps, pi = merge_topk(p_sc, p_ids, larger)
        assert np.array_equal(pi.cpu().numpy(), mi)
        assert np.array_equal(ps.cpu().numpy(), ms, equal_nan=True)


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (1M x 384 fp32, top-10): size-independent properties plus a
    direct comparison with oracle.c on a few queries."""
    import torch
    from archi_b200.store import NativeStore
    n, d = 1_000_000, 384
    g = torch.Generator(device="cuda").manual_seed(1234 + 2000)
    x = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)
    x = x / x.norm(dim=1, keepdim=True)
    s = NativeStore(d, "cosine", "f32", capacity_rows=n)
    s.append(x)
    # (1) every probed row finds itself first with similarity 1
    probe = torch.tensor([0, 1, 31, 32, 499_999, 999_998, 999_999], device="cuda")
    sc, ids = s.search(x[probe], 10)
    torch.cuda.synchronize()
    assert (ids[:, 0] == probe).all()
    assert torch.allclose(sc[:, 0], torch.ones(7, device="cuda"), atol=2e-6)
    assert (sc[:, :-1] >= sc[:, 1:]).all()
    # (2) batch-of-8 and one-by-one give identical answers (different kernels instantiations)
    gq = torch.Generator(device="cuda").manual_seed(4321 + 2000)
    q = torch.randn((8, d), generator=gq, device="cuda")
    q = q / q.norm(dim=1, keepdim=True)
    sc8, id8 = s.search(q, 10)
    for i in range(8):
        sc1, id1 = s.search(q[i:i + 1], 10)
        assert torch.equal(id1[0], id8[i]) and torch.allclose(sc1[0], sc8[i], rtol=1e-6, atol=1e-7)
    # (3) oracle.c on the same stored values, 4 queries
    xc = x.cpu().numpy()
    dq, iq = orc.c_scan_topk("cosine", xc, q[:4].cpu().numpy(), 10, nthreads=4)
    for i in range(4):
        assert orc.same_topk_up_to_ties(id8[i].cpu().tolist(), iq[i], dq[i], rel_tol=2e-6, abs_tol=1e-7)
    assert np.allclose(sc8[:4].cpu().numpy(), 1.0 - dq, rtol=REL_F32, atol=1e-6)
    s.close()


@pytest.mark.parametrize("storage", ["f32", "bf16"])
@pytest.mark.parametrize("nq", [2, 3, 5, 8, 11])
def test_streaming_path_multi_query_kernels(storage, nq):
    """ARCHI_PATH_AUTO sends batches >= 2 to the tensor path; the multi-query streaming kernels
    (QB = 2 / 4 / 8) still serve hybrid search and proof fallbacks, so they are pinned here."""
    rng = np.random.default_rng(100 + nq)
    corpus = unit_rows(rng, 15013, 200)
    queries = unit_rows(rng, nq, 200)
    for metric in ("cosine", "l2"):
        s = make_store(corpus, metric, storage)
        scores, ids = s.search(queries, 10, path=1)
        assert s.last_stats().path == 1
        check_against_truth(metric, stored_values(corpus, storage), queries, 10, scores, ids,
                            REL_BF16 if storage == "bf16" else REL_F32)
        s.close()
