"""CPU tests of the oracle: the reference's own known answers (score conventions), agreement of
oracle.c with the fp64 numpy truth, and the committed golden fixtures."""
import os

import numpy as np
import pytest

from oracle import oracle as orc


# ---- known answers taken from the reference's tests (tests/unit/test_postgres_vectorstore.py) ------
def test_operator_map_matches_reference():
    # :94-128 -- cosine <=>, l2 <->, inner_product <#>
    assert orc.DISTANCE_OPS == {"cosine": "<=>", "l2": "<->", "inner_product": "<#>"}


def test_score_is_one_minus_distance_for_cosine():
    # :196,210,462,474 -- distance 0.1 -> score 0.9
    assert orc.score_from_distance("cosine", 0.1) == pytest.approx(0.9)
    assert orc.clib().orc_score_from_distance(0, 0.1) == pytest.approx(0.9)
    # l2 / inner_product return the raw distance (postgres_vectorstore.py:361)
    assert orc.score_from_distance("l2", 0.25) == 0.25
    assert orc.score_from_distance("inner_product", -0.75) == -0.75


@pytest.mark.parametrize("sem,bm,ws,wb,expected", [
    (0.85, 0.9, 0.7, 0.3, 0.865),   # :259-261
    (0.80, 0.7, 0.4, 0.6, 0.74),    # :304-306
    (0.95, 0.2, 0.7, 0.3, 0.725),   # :352-354
    (0.70, 0.8, 0.7, 0.3, 0.73),    # :364-366
])
def test_hybrid_combination_known_answers(sem, bm, ws, wb, expected):
    # a 1-d corpus whose cosine similarity to the query is +1; fold the wanted semantic score in by
    # using inner_product (semantic = 1 - (-a.b) = 1 + a.b)
    corpus = np.array([[sem - 1.0]], dtype=np.float32)
    q = np.array([1.0], dtype=np.float32)
    comb, ids = orc.exact_hybrid_topk("inner_product", corpus, q, np.array([bm]), ws, wb, 1)
    assert ids.tolist() == [0]
    assert comb[0] == pytest.approx(expected, abs=1e-6)
    cc, ic = orc.c_hybrid_topk("inner_product", corpus, q, np.array([bm]), ws, wb, 1)
    assert cc[0] == pytest.approx(expected, abs=1e-6)


def test_hybrid_ordering_known_answer():
    # :352-366 -- (0.95 sem, 0.2 bm25) = 0.725 ranks below (0.7 sem, 0.8 bm25) = 0.73 at 0.7/0.3
    corpus = np.array([[0.95 - 1.0], [0.70 - 1.0]], dtype=np.float32)
    q = np.array([1.0], dtype=np.float32)
    comb, ids = orc.exact_hybrid_topk("inner_product", corpus, q, np.array([0.2, 0.8]), 0.7, 0.3, 2)
    assert ids.tolist() == [1, 0]


def test_hybrid_coalesce_null_bm25():
    corpus = np.eye(3, dtype=np.float32)
    q = np.array([1, 0, 0], dtype=np.float32)
    bm25 = np.array([np.nan, 5.0, np.nan])
    comb, ids = orc.exact_hybrid_topk("cosine", corpus, q, bm25, 0.5, 0.5, 3)
    assert ids.tolist() == [1, 0, 2]
    assert comb.tolist() == pytest.approx([2.5, 0.5, 0.0])


def test_mock_embedding_fixture_vector():
    # the reference's mock embedding [0.1,0.2,0.3]*128 (:48-49): cosine distance to itself is 0
    v = np.array([0.1, 0.2, 0.3] * 128, dtype=np.float32)
    d = orc.distances_f64("cosine", v[None, :], v[None, :])
    assert d[0, 0] == pytest.approx(0.0, abs=1e-12)
    assert orc.clib().orc_distance_f32(0, 384, v.ctypes.data, v.ctypes.data) == pytest.approx(0.0, abs=1e-6)


def test_smoke_property_ascending_and_min_k_rows():
    # tests/smoke/test_integration.py:516-534: 5 random 384-d vectors, LIMIT 3 -> 3 rows ascending
    rng = np.random.default_rng(5)
    corpus = rng.random((5, 384)).astype(np.float32)
    q = rng.random(384).astype(np.float32)
    d, i = orc.c_scan_topk("cosine", corpus, q, 3)
    assert (i[0] >= 0).all() and len(set(i[0].tolist())) == 3
    assert (np.diff(d[0]) >= 0).all()
    d, i = orc.c_scan_topk("cosine", corpus, q, 10)       # k > N: N rows then padding
    assert (i[0, :5] >= 0).all() and (i[0, 5:] == -1).all()


# ---- oracle.c (float accumulators) against the fp64 truth ---------------------------------------------
@pytest.mark.parametrize("metric", list(orc.METRICS))
def test_c_scan_matches_fp64_truth(metric):
    rng = np.random.default_rng(11)
    corpus = rng.standard_normal((4000, 128)).astype(np.float32)
    queries = rng.standard_normal((4, 128)).astype(np.float32)
    d, i = orc.exact_topk(metric, corpus, queries, 16)
    dc, ic = orc.c_scan_topk(metric, corpus, queries, 16, nthreads=3)
    for q in range(4):
        assert orc.same_topk_up_to_ties(ic[q], i[q], d[q], rel_tol=1e-6)
    assert np.allclose(d, dc, rtol=1e-5, atol=1e-5)


def test_mask_and_ties():
    corpus = np.tile(np.array([[1.0, 0.0]], dtype=np.float32), (8, 1))   # 8 identical rows
    q = np.array([1.0, 0.0], dtype=np.float32)
    d, i = orc.c_scan_topk("cosine", corpus, q, 3)
    assert i[0].tolist() == [0, 1, 2]                  # ties -> lower id first
    mask = np.array([0, 0, 1, 1, 0, 1, 1, 1], dtype=bool)
    d, i = orc.c_scan_topk("cosine", corpus, q, 3, mask=mask)
    assert i[0].tolist() == [2, 3, 5]
    d2, i2 = orc.exact_topk("cosine", corpus, q, 3, mask=mask)
    assert i2[0].tolist() == [2, 3, 5]


def test_bf16_roundtrip_helpers():
    x = np.array([1.0, 0.1, -3.14159, 1e-3, 65504.0], dtype=np.float32)
    b = orc.f32_to_bf16_bits(x)
    y = orc.bf16_bits_to_f32(b)
    assert np.allclose(x, y, rtol=2 ** -8)
    assert orc.bf16_bits_to_f32(orc.f32_to_bf16_bits(y)).tolist() == y.tolist()   # idempotent


def test_pool_normalize_c_matches_numpy_and_is_unit_norm():
    rng = np.random.default_rng(3)
    h = rng.standard_normal((5, 17, 32)).astype(np.float32)
    m = (rng.random((5, 17)) < 0.7).astype(np.int64)
    m[0] = 1
    a = orc.pool_normalize(h, m)
    c = orc.c_pool_normalize(h, m)
    assert np.allclose(a, c, atol=2e-6)
    nz = m.sum(1) > 0
    assert np.allclose(np.linalg.norm(a[nz], axis=1), 1.0, atol=1e-9)


def test_bm25_restatement_basics():
    docs = [orc.tokenize(t) for t in ["the cat sat", "the dog sat on the cat", "quantum chromodynamics", ""]]
    s = orc.bm25_scores(docs, orc.tokenize("cat"))
    assert np.isnan(s[2]) and np.isnan(s[3])           # no shared term -> NULL
    assert s[0] > s[1] > 0                              # shorter doc with the same tf ranks higher
    assert (orc.bm25_scores(docs, ["cat"], sign=-1.0)[:2] < 0).all()


def test_character_text_split_default_config():
    text = "\n\n".join(["a" * 400, "b" * 400, "c" * 400, "d" * 1500, "e" * 10])
    chunks = orc.character_text_split(text, 1000, 0)
    assert chunks[0] == "a" * 400 + "\n\n" + "b" * 400
    assert "d" * 1500 in chunks                         # oversize piece kept whole
    assert all(c for c in chunks)


# ---- committed golden fixtures --------------------------------------------------------------------------
def test_golden_search_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "search_2048x96.npz"))
    for metric in orc.METRICS:
        d, i = orc.exact_topk(metric, g["corpus"], g["queries"], 10)
        assert (i == g[f"{metric}_ids"]).all()
        assert np.allclose(d, g[f"{metric}_dist"], rtol=1e-12, atol=1e-12)
        dc, ic = orc.c_scan_topk(metric, g["corpus"], g["queries"], 10)
        assert (ic == g[f"{metric}_ids"]).all()
    stored = orc.bf16_bits_to_f32(g["bf16_bits"])
    d, i = orc.exact_topk("cosine", stored, g["queries"], 10)
    assert (i == g["bf16_cosine_ids"]).all()
    dc, ic = orc.c_scan_topk("cosine", g["bf16_bits"], g["queries"], 10, corpus_is_bf16=True)
    assert (ic == g["bf16_cosine_ids"]).all()


def test_golden_hybrid_and_pool_fixtures(golden_dir):
    h = np.load(os.path.join(golden_dir, "hybrid_512x64.npz"))
    for metric in orc.METRICS:
        for ws, wb in ((0.7, 0.3), (0.4, 0.6)):
            cc, ic = orc.c_hybrid_topk(metric, h["corpus"], h["query"], h["bm25"], ws, wb, 8)
            assert (ic == h[f"{metric}_{ws}_{wb}_ids"]).all()
            assert np.allclose(cc, h[f"{metric}_{ws}_{wb}_combined"], atol=1e-5)
    p = np.load(os.path.join(golden_dir, "pool_6x24x64.npz"))
    assert np.allclose(orc.c_pool_normalize(p["hidden"], p["mask"]), p["pooled"], atol=2e-6)


def test_hnsw_restatement_recall_and_truncation():
    """The reference's default index (HNSW m=16, ef_construction=64, ef_search=40), restated: near-exact
    on low-intrinsic-dimension rows, and an index scan never returns more than ef_search rows."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3000, 16)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((40, 16)).astype(np.float32)
    idx = orc.HnswIndex(x)
    _, exact = orc.exact_topk("cosine", x, q, 10)
    _, approx = idx.search(q, 10, ef_search=40)
    assert orc.recall_at_k(approx, exact) > 0.9
    _, approx = idx.search(q, 100, ef_search=40)
    assert (approx[:, 40:] == -1).all() and (approx[:, :40] >= 0).all()
    idx.close()
