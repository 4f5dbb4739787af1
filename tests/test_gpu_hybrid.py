"""GPU parity of the posting-list hybrid search (archi_hybrid_search_terms, csrc/hybrid.cu) against the oracle's
hybrid (postgres_vectorstore.py:435-457 restated: combined = (1 - distance) * w_sem + COALESCE(bm25, 0) * w_bm25,
ORDER BY combined DESC LIMIT k) with the restated BM25 as the lexical score -- both the posting-list path and the
dense-vector fallback, single queries and batches, filters and tombstones, all three metrics."""
import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_search import make_store, unit_rows

pytestmark = pytest.mark.gpu

VOCAB = 400


def _corpus(rng, n, dim, doc_len=12, scale=False):
    x = unit_rows(rng, n, dim)
    if scale:
        x = (x * rng.uniform(0.5, 1.5, size=(n, 1))).astype(np.float32)
    # zipf-ish term ids: a few very frequent terms (dense regime) and a long tail (sparse regime)
    tokens = np.minimum((rng.zipf(1.4, size=(n, doc_len)) - 1), VOCAB - 1).astype(np.int64)
    return x, tokens


def _truth(metric, stored, tokens, q_vec, q_terms, ws, wb, k, mask=None):
    docs = [[str(t) for t in row] for row in tokens]
    live_docs = docs if mask is None else docs
    bm = orc.bm25_scores(live_docs, [str(t) for t in q_terms])
    return orc.exact_hybrid_topk(metric, stored, q_vec, bm, ws, wb, k, mask=mask)


def _check(metric, stored, tokens, queries, q_terms, ws, wb, k, scores, ids, mask=None, bm_docs=None):
    for i in range(queries.shape[0]):
        docs = bm_docs if bm_docs is not None else [[str(t) for t in row] for row in tokens]
        bm = orc.bm25_scores(docs, [str(t) for t in q_terms[i]])
        comb, cid = orc.exact_hybrid_topk(metric, stored, queries[i], bm, ws, wb, k, mask=mask)
        kk = len(cid)
        got_i, got_s = ids[i, :kk], scores[i, :kk]
        assert (ids[i, kk:] == -1).all()
        assert np.allclose(got_s, comb, rtol=2e-5, atol=3e-6), (i, got_s, comb)
        if got_i.tolist() != cid.tolist():
            # only near-ties may permute
            kth = comb[-1]
            strict = {int(c) for c, s in zip(cid, comb) if s > kth + 1e-5 * max(1.0, abs(kth))}
            assert strict.issubset(set(got_i.tolist())), (i, got_i, cid)


@pytest.mark.parametrize("metric", ["cosine", "l2", "inner_product"])
@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_posting_list_hybrid_matches_oracle(metric, storage):
    import torch
    from archi_b200.bm25 import LexicalIndex
    rng = np.random.default_rng(31)
    n, dim = 20011, 96
    corpus, tokens = _corpus(rng, n, dim, scale=(metric != "cosine"))
    s = make_store(corpus, metric, storage)
    stored = orc.bf16_bits_to_f32(orc.f32_to_bf16_bits(corpus)) if storage == "bf16" else corpus
    lex = LexicalIndex(0)
    lex.add_token_matrix(tokens)
    queries = unit_rows(rng, 40, dim)
    # rare terms only -> posting-list path; term 0 / 1 are in most rows -> dense-vector path
    rare = [rng.integers(60, VOCAB, size=3) for _ in range(40)]
    for k, ws, wb in ((5, 0.7, 0.3), (10, 0.4, 0.6), (40, 0.5, 0.5)):
        sc, ids = s.hybrid_search_terms(lex, rare, queries, k, ws, wb)
        assert s.last_hybrid_path == "posting-lists"
        _check(metric, stored, tokens, queries, rare, ws, wb, k, sc, ids)
        sc1, id1 = s.hybrid_search_terms(lex, rare[:1], queries[:1], k, ws, wb)        # one query: streaming dense part
        assert np.array_equal(id1[0], ids[0]) and np.allclose(sc1[0], sc[0], rtol=1e-6, atol=1e-6)
    common = [np.array([0, 1, int(rng.integers(60, VOCAB))]) for _ in range(6)]
    sc, ids = s.hybrid_search_terms(lex, common, queries[:6], 8, 0.4, 0.6)
    assert s.last_hybrid_path == "dense-vector"
    _check(metric, stored, tokens, queries[:6], common, 0.4, 0.6, 8, sc, ids)
    # device tensors in, device tensors out; queries without any known term behave like a dense search
    none = [np.array([VOCAB + 5]) for _ in range(3)]
    sc_d, id_d = s.hybrid_search_terms(lex, none, torch.from_numpy(queries[:3]).cuda(), 5, 0.7, 0.3)
    dsc, did = s.search(queries[:3], 5)
    torch.cuda.synchronize()
    assert np.array_equal(id_d.cpu().numpy(), did)
    sem = dsc if metric == "cosine" else 1.0 - dsc
    assert np.allclose(sc_d.cpu().numpy(), 0.7 * sem, rtol=1e-6, atol=1e-6)
    s.close()


def test_posting_list_hybrid_with_filter_tombstones_and_repeated_terms():
    import torch
    from archi_b200.bm25 import LexicalIndex
    rng = np.random.default_rng(32)
    n, dim = 9001, 64
    corpus, tokens = _corpus(rng, n, dim)
    s = make_store(corpus)
    lex = LexicalIndex(0)
    lex.add_token_matrix(tokens)
    queries = unit_rows(rng, 20, dim)
    terms = [np.array([int(t)] * 2 + [int(u)]) for t, u in zip(rng.integers(60, VOCAB, 20), rng.integers(60, VOCAB, 20))]
    keep = rng.random(n) < 0.5
    fm = torch.from_numpy(np.concatenate([orc.pack_mask(keep), np.zeros(1, np.uint32)]).view(np.int32).copy()).cuda()
    sc, ids = s.hybrid_search_terms(lex, terms, queries, 7, 0.4, 0.6, filter_mask=fm)
    assert s.last_hybrid_path == "posting-lists"
    _check("cosine", corpus, tokens, queries, terms, 0.4, 0.6, 7, sc, ids, mask=keep)
    # tombstones: the rows leave the store AND the lexical statistics (as DELETE does in the reference)
    gone = ids[0, :3].tolist()
    s.delete_rows(gone)
    lex.delete_rows(gone)
    alive = np.ones(n, dtype=bool)
    alive[gone] = False
    docs = [[str(t) for t in row] for row in tokens]
    live_docs = [d for d, a in zip(docs, alive) if a]
    sc, ids = s.hybrid_search_terms(lex, terms, queries, 7, 0.4, 0.6)
    for i in range(20):
        bm_live = orc.bm25_scores(live_docs, [str(t) for t in terms[i]])
        bm = np.full(n, np.nan)
        bm[alive] = bm_live
        comb, cid = orc.exact_hybrid_topk("cosine", corpus, queries[i], bm, 0.4, 0.6, 7, mask=alive)
        assert np.allclose(sc[i], comb, rtol=2e-5, atol=3e-6) and set(ids[i].tolist()) == set(cid.tolist())
    # the accumulator is left clean: the same call twice gives the same bits
    sc2, ids2 = s.hybrid_search_terms(lex, terms, queries, 7, 0.4, 0.6)
    assert np.array_equal(sc, sc2) and np.array_equal(ids, ids2)
    s.close()


def test_posting_list_hybrid_at_1m_rows_sampled():
    """Config 5's scale: 1M x 384 fp32, 24-term documents; a batch of 64 and single queries, sampled against oracle.c."""
    import torch
    from archi_b200.bm25 import LexicalIndex
    from archi_b200.store import NativeStore
    n, dim = 1_000_000, 384
    g = torch.Generator(device="cuda").manual_seed(55)
    s = NativeStore(dim, "cosine", "f32", capacity_rows=n)
    host = np.empty((n, dim), dtype=np.float32)
    for r0 in range(0, n, 250_000):
        x = torch.randn((250_000, dim), generator=g, device="cuda")
        x = x / x.norm(dim=1, keepdim=True)
        s.append(x)
        host[r0:r0 + 250_000] = x.cpu().numpy()
    rng = np.random.default_rng(5)
    tokens = (rng.zipf(1.3, size=(n, 24)) % 50_000).astype(np.int64)
    lex = LexicalIndex(0)
    lex.add_token_matrix(tokens)
    q = torch.randn((64, dim), generator=g, device="cuda")
    q = q / q.norm(dim=1, keepdim=True)
    terms = [((rng.zipf(1.3, size=3) % 49_900) + 100).astype(np.int64) for _ in range(64)]
    sc, ids = s.hybrid_search_terms(lex, terms, q, 5, 0.4, 0.6)
    assert s.last_hybrid_path == "posting-lists"
    sc, ids, qh = sc.cpu().numpy(), ids.cpu().numpy(), q.cpu().numpy()
    for i in range(0, 64, 8):
        bm = lex.score(terms[i]).cpu().numpy().astype(np.float64)          # dense BM25 from the accumulate kernel
        comb, cid = orc.c_hybrid_topk("cosine", host, qh[i], np.where(bm != 0, bm, np.nan), 0.4, 0.6, 5)
        assert set(ids[i].tolist()) == set(cid.tolist()), (i, ids[i], cid)
        assert np.allclose(np.sort(sc[i]), np.sort(comb), rtol=1e-5, atol=2e-6)
        s1, i1 = s.hybrid_search_terms(lex, [terms[i]], qh[i:i + 1], 5, 0.4, 0.6)
        assert np.array_equal(i1[0], ids[i])
    s.close()
