"""GPU parity tests of the tensor-core search path (tcgen05/TMEM coarse scorer + exact rescoring).
Same bar as the streaming path: ids identical to the fp64 exact search over the stored values up to
tie order, scores within 1e-5 relative (fp32 storage) / 2e-3 (bf16 storage) -- the coarse pass only
nominates candidates; every returned score is recomputed in fp32 from the stored row."""
import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_search import REL_BF16, REL_F32, check_against_truth, make_store, stored_values, unit_rows

pytestmark = pytest.mark.gpu

TENSOR = 2  # ARCHI_PATH_TENSOR


def run_tensor(s, queries, k, **kw):
    scores, ids = s.search(queries, k, path=TENSOR, **kw)
    st = s.last_stats()
    assert st.path == TENSOR
    return scores, ids, st


@pytest.mark.parametrize("storage", ["f32", "bf16"])
@pytest.mark.parametrize("metric", ["cosine", "l2", "inner_product"])
@pytest.mark.parametrize("nq", [1, 17, 128, 300])
def test_tensor_path_matches_oracle(metric, storage, nq):
    rng = np.random.default_rng(4321 + nq)
    corpus = (unit_rows(rng, 30011, 384) * rng.uniform(0.5, 1.5, size=(30011, 1))).astype(np.float32)
    queries = unit_rows(rng, nq, 384) * rng.uniform(0.5, 2.0, size=(nq, 1)).astype(np.float32)
    s = make_store(corpus, metric, storage)
    scores, ids, st = run_tensor(s, queries.astype(np.float32), 10)
    check_against_truth(metric, stored_values(corpus, storage), queries, 10, scores, ids,
                        REL_BF16 if storage == "bf16" else REL_F32)
    assert st.unverified_queries <= max(1, nq // 20)      # the proof almost always holds
    s.close()


@pytest.mark.parametrize("dim", [8, 30, 100, 257, 768, 1024])
@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_tensor_path_dims(dim, storage):
    rng = np.random.default_rng(dim)
    corpus = rng.standard_normal((5000, dim)).astype(np.float32)
    queries = rng.standard_normal((40, dim)).astype(np.float32)
    s = make_store(corpus, "cosine", storage)
    scores, ids, st = run_tensor(s, queries, 5)
    check_against_truth("cosine", stored_values(corpus, storage), queries, 5, scores, ids,
                        REL_BF16 if storage == "bf16" else REL_F32)
    s.close()


@pytest.mark.parametrize("k", [1, 10, 32, 100, 128])
def test_tensor_path_k_values(k):
    rng = np.random.default_rng(k)
    corpus = unit_rows(rng, 40000, 128)
    queries = unit_rows(rng, 33, 128)
    for storage in ("f32", "bf16"):
        s = make_store(corpus, "cosine", storage)
        scores, ids, st = run_tensor(s, queries, k)
        check_against_truth("cosine", stored_values(corpus, storage), queries, k, scores, ids,
                            REL_BF16 if storage == "bf16" else REL_F32)
        s.close()


def test_tensor_path_small_corpus_and_k_larger_than_rows():
    rng = np.random.default_rng(1)
    corpus = unit_rows(rng, 37, 64)
    queries = unit_rows(rng, 20, 64)
    s = make_store(corpus)
    scores, ids, st = run_tensor(s, queries, 50)
    assert (ids[:, :37] >= 0).all() and (ids[:, 37:] == -1).all() and np.isnan(scores[:, 37:]).all()
    check_against_truth("cosine", corpus, queries, 50, scores, ids, REL_F32)
    s.close()


def test_tensor_path_filter_and_tombstones():
    import torch
    rng = np.random.default_rng(99)
    corpus = unit_rows(rng, 25013, 96)
    queries = unit_rows(rng, 70, 96)
    s = make_store(corpus, "cosine", "bf16")
    stored = stored_values(corpus, "bf16")
    keep = rng.random(25013) < 0.25
    words = np.concatenate([orc.pack_mask(keep), np.zeros(1, np.uint32)])
    fm = torch.from_numpy(words.view(np.int32).copy()).cuda()
    scores, ids, _ = run_tensor(s, queries, 10, filter_mask=fm)
    check_against_truth("cosine", stored, queries, 10, scores, ids, REL_BF16, mask=keep)
    _, top, _ = run_tensor(s, queries[:1], 5)
    s.delete_rows(top[0])
    alive = np.ones(25013, dtype=bool)
    alive[top[0]] = False
    scores, ids, _ = run_tensor(s, queries, 10)               # cached row constants must be rebuilt
    check_against_truth("cosine", stored, queries, 10, scores, ids, REL_BF16, mask=alive)
    scores, ids, _ = run_tensor(s, queries, 10, filter_mask=fm)
    check_against_truth("cosine", stored, queries, 10, scores, ids, REL_BF16, mask=alive & keep)
    scores, ids, _ = run_tensor(s, queries, 10)               # and again without the filter
    check_against_truth("cosine", stored, queries, 10, scores, ids, REL_BF16, mask=alive)
    s.close()


def test_tensor_path_duplicates_and_failed_proofs_fall_back_to_exact():
    """Many identical rows tie exactly (compaction must not overflow) and a corpus whose scores are
    all within eps of each other cannot be proven by the coarse pass: those queries are re-run on the
    exact streaming path and still come back right."""
    rng = np.random.default_rng(3)
    base = unit_rows(rng, 40, 64)
    corpus = np.repeat(base, 500, axis=0)                      # 20000 rows, 40 distinct
    queries = unit_rows(rng, 12, 64)
    s = make_store(corpus)
    scores, ids, st = run_tensor(s, queries, 10)
    d_true, i_true = orc.exact_topk("cosine", corpus, queries, 10)
    assert np.allclose(scores, 1.0 - d_true, rtol=REL_F32, atol=1e-6)
    for q in range(12):                                        # any 10 of the 500 tied copies are right
        assert len(set(ids[q].tolist())) == 10
        best = int(i_true[q, 0]) // 500
        assert all(i // 500 == best for i in ids[q].tolist())
    s.close()
    # near-identical rows: exact differences of ~1e-5 are far below the tf32 error bound
    center = unit_rows(rng, 1, 96)
    corpus = (center + 1e-5 * rng.standard_normal((6000, 96))).astype(np.float32)
    s = make_store(corpus, "inner_product")
    qq = unit_rows(rng, 12, 96)
    scores, ids, st = run_tensor(s, qq, 10)
    assert st.unverified_queries > 0
    check_against_truth("inner_product", corpus, qq, 10, scores, ids, REL_F32)
    s.close()


def test_failed_proofs_are_rescued_on_the_device_without_a_host_round_trip():
    """Device outputs: nothing synchronises inside archi_search, the queries whose proof failed are re-scanned
    by launches driven from a device-side list.  Every query of a batch <= 256 can be rescued."""
    import torch
    rng = np.random.default_rng(5)
    center = unit_rows(rng, 1, 96)
    corpus = (center + 1e-5 * rng.standard_normal((6000, 96))).astype(np.float32)
    s = make_store(corpus, "inner_product")
    qq = unit_rows(rng, 200, 96)
    sc, ids = s.search(torch.from_numpy(qq).cuda(), 10, path=TENSOR)
    st = s.last_stats()                                        # synchronises with the search
    assert st.unverified_queries > 100 and st.unproven_queries == 0
    check_against_truth("inner_product", corpus, qq, 10, sc.cpu().numpy(), ids.cpu().numpy(), REL_F32)
    s.close()


def test_more_failed_proofs_than_the_rescue_list_holds():
    """Batches > 256 whose proofs (almost) all fail: with host outputs the call re-scans every flagged query
    before it returns (still exact); with device outputs the queries beyond the 256 rescued ones come back as
    id -1 / score NaN -- never as an unproven row -- and are counted in unproven_queries."""
    import torch
    rng = np.random.default_rng(6)
    center = unit_rows(rng, 1, 64)
    corpus = (center + 1e-5 * rng.standard_normal((5000, 64))).astype(np.float32)
    qq = unit_rows(rng, 300, 64)
    s = make_store(corpus, "inner_product")
    scores, ids = s.search(qq, 10, path=TENSOR)                # host buffers
    assert s.last_stats().unverified_queries > 256
    check_against_truth("inner_product", corpus, qq, 10, scores, ids, REL_F32)
    sc_d, id_d = s.search(torch.from_numpy(qq).cuda(), 10, path=TENSOR)
    st = s.last_stats()
    sc_d, id_d = sc_d.cpu().numpy(), id_d.cpu().numpy()
    poisoned = (id_d[:, 0] == -1)
    assert st.unproven_queries == poisoned.sum() == st.unverified_queries - 256
    assert (id_d[poisoned] == -1).all() and np.isnan(sc_d[poisoned]).all()
    good = ~poisoned
    assert np.array_equal(id_d[good], ids[good]) and np.array_equal(sc_d[good], scores[good])
    s.close()


def test_searches_on_different_streams_are_ordered_on_the_device():
    """The handle's scratch is shared: a search on another stream waits for the previous call's work."""
    import torch
    rng = np.random.default_rng(7)
    corpus = unit_rows(rng, 50000, 128)
    s = make_store(corpus)
    qa, qb = unit_rows(rng, 64, 128), unit_rows(rng, 64, 128)
    want_a, want_b = s.search(qa, 10), s.search(qb, 10)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    da, db = torch.from_numpy(qa).cuda(), torch.from_numpy(qb).cuda()
    torch.cuda.synchronize()
    outs = []
    for rep in range(6):
        with torch.cuda.stream(s1):
            outs.append(("a", s.search(da, 10)))
        with torch.cuda.stream(s2):
            outs.append(("b", s.search(db, 10)))
    torch.cuda.synchronize()
    for which, (sc, ids) in outs:
        want = want_a if which == "a" else want_b
        assert np.array_equal(ids.cpu().numpy(), want[1]) and np.array_equal(sc.cpu().numpy(), want[0])
    s.close()


def test_auto_path_switches_by_batch_size():
    rng = np.random.default_rng(8)
    corpus = unit_rows(rng, 9000, 64)
    s = make_store(corpus)
    s.search(unit_rows(rng, 1, 64), 5)
    assert s.last_stats().path == 1                            # streaming for a single query
    sc_t, id_t = s.search(unit_rows(rng, 64, 64), 5)
    assert s.last_stats().path == TENSOR
    # hybrid and k > 128 stay on the streaming path
    s.search(unit_rows(rng, 64, 64), 200)
    assert s.last_stats().path == 1
    s.close()


def test_streaming_and_tensor_paths_agree_at_full_size():
    """Config 2 at full size (1M x 384 fp32, top-10, 256 queries): BOTH paths against oracle.c on every
    query, tie-aware and exact (every id outside the oracle's list is re-scored in fp64 and must tie the
    k-th distance), and the two CUDA paths bit-compatible with each other."""
    import os
    import torch
    from archi_b200.store import NativeStore
    n, d = 1_000_000, 384
    g = torch.Generator(device="cuda").manual_seed(77)
    s = NativeStore(d, "cosine", "f32", capacity_rows=n)
    host = np.empty((n, d), dtype=np.float32)
    for r0 in range(0, n, 250_000):
        x = torch.randn((250_000, d), generator=g, device="cuda")
        x = x / x.norm(dim=1, keepdim=True)
        s.append(x)
        host[r0:r0 + 250_000] = x.cpu().numpy()
    q = torch.randn((256, d), generator=g, device="cuda")
    q = q / q.norm(dim=1, keepdim=True)
    sc_t, id_t = s.search(q, 10, path=TENSOR)
    stats = s.last_stats()
    sc_s, id_s = s.search(q, 10, path=1)
    torch.cuda.synchronize()
    assert stats.path == TENSOR and stats.unverified_queries == 0
    qh = q.cpu().numpy()
    d_true, i_true = orc.c_scan_topk("cosine", host, qh, 10, nthreads=os.cpu_count() or 4)
    for sc, ids in ((sc_t, id_t), (sc_s, id_s)):
        fails = orc.verify_topk("cosine", host, qh, 10, ids.cpu().numpy(), sc.cpu().numpy(), REL_F32, d_true, i_true)
        assert not fails, fails[:5]
    # the two CUDA paths: same fp32 scores rank by rank; ids equal wherever the scores are not tied
    assert torch.allclose(sc_t, sc_s, rtol=1e-6, atol=1e-7)
    diff = id_t != id_s
    if diff.any():
        qi, ri = torch.nonzero(diff, as_tuple=True)
        for a, b in zip(qi.tolist(), ri.tolist()):
            near = torch.isclose(sc_s[a], sc_s[a, b], rtol=2e-6, atol=1e-7).sum().item()
            assert near >= 2, (a, b, id_t[a].tolist(), id_s[a].tolist())
    s.close()
