"""CPU tests of the lexical index's host side (archi_b200/bm25.py + hostsrc/text_index.c): the posting
lists it hands to archi_bm25_accumulate, scored here with numpy by the kernel's formula, must reproduce
the restated BM25 -- on the C helper path and on the Python path alike."""
import random
import time

import numpy as np
import pytest

from archi_b200 import bm25 as B
from oracle import oracle as orc


def bm25_from_csr(ix, query):
    """What archi_bm25_accumulate computes (include/archi_b200.h), on the host CSR."""
    for m in ix.table.members:       # the statistics are table-wide: every index of the table must be current
        m._term_keys, m._df, m._post_ptr, _, _, _, m._n_live, m._avgdl = m._host_csr()
    term_keys, df, post_ptr, doc_ids, tfs, dl, n_live, avgdl = ix._host_csr()
    avgdl = ix.avgdl()
    out = np.zeros(len(ix), dtype=np.float64)
    touched = np.zeros(len(ix), dtype=bool)
    for t in ix.query_terms(query):
        if df[t] == 0:
            continue
        sl = slice(post_ptr[t], post_ptr[t + 1])
        d, tf = doc_ids[sl], tfs[sl].astype(np.float64)
        out[d] += ix.idf(t) * tf * (ix.k1 + 1) / (tf + ix.k1 * (1 - ix.b + ix.b * dl[d] / avgdl)) * ix.sign
        touched[d] = True
    return np.where(touched, out, np.nan)


def corpus(rng, n, unicode_every=0):
    words = ["alpha", "Beta", "GAMMA", "delta9", "x1", "the", "muon", "Higgs", "boson", "42", "e", "quark"]
    docs = []
    for i in range(n):
        toks = [rng.choice(words) for _ in range(rng.randint(0, 40))]
        text = rng.choice([" ", ", ", "\n", " - ", "/"]).join(toks)
        if unicode_every and i % unicode_every == 0:
            text += " Übergröße naïve ΣΩ K İstanbul " + rng.choice(words)       # Kelvin sign, dotted capital I
        docs.append(text)
    return docs


@pytest.mark.parametrize("fast", [True, False])
def test_postings_reproduce_restated_bm25(fast, monkeypatch):
    if fast and B._text_lib() is None:
        pytest.skip("libarchi_text.so not built")
    if not fast:
        monkeypatch.setattr(B, "_text_lib", lambda: None)
    rng = random.Random(11)
    docs = corpus(rng, 300, unicode_every=7)
    ix = B.LexicalIndex(0)
    assert ix._fast is fast
    ix.add_texts(docs[:100])
    ix.add_texts(docs[100:])
    ix.delete_rows([3, 50, 299])
    tokenised = [orc.tokenize(t) for t in docs]
    for i in (3, 50, 299):
        tokenised[i] = None                                                      # deleted: no postings, not counted
    live_docs = [t for t in tokenised if t is not None]
    for query in ("higgs boson", "the the muon", "nothing matches zzz", "K 42 übergröße", "E"):
        want_live = orc.bm25_scores(live_docs, orc.tokenize(query))
        want = np.full(len(docs), np.nan)
        want[[i for i, t in enumerate(tokenised) if t is not None]] = want_live
        got = bm25_from_csr(ix, query)
        assert np.array_equal(np.isnan(got), np.isnan(want)), query
        assert np.allclose(got[~np.isnan(got)], want[~np.isnan(want)], rtol=1e-6), query


def test_helper_tokenises_like_the_python_tokenizer():
    lib = B._text_lib()
    if lib is None:
        pytest.skip("libarchi_text.so not built")
    rng = random.Random(2)
    docs = corpus(rng, 200, unicode_every=3) + ["", "   ", "A", "a-b_c.d", "\x00x\x7fy", "ß"]
    ix = B.LexicalIndex(0)
    ix.add_texts(docs)
    keys, tfs, counts, lens = ix._batches[0]
    at = 0
    for d, n_pairs, n_tok in zip(docs, counts, lens):
        toks = B.default_tokenize(d)
        assert n_tok == len(toks)
        want = {}
        for t in toks:
            k = lib.archi_text_term_key(t.encode("ascii"), len(t))
            want[k] = want.get(k, 0) + 1
        got = dict(zip(keys[at:at + n_pairs].tolist(), tfs[at:at + n_pairs].tolist()))
        assert got == want
        assert np.all(np.diff(keys[at:at + n_pairs].astype(np.float64)) > 0) or n_pairs < 2   # ascending, distinct
        at += n_pairs
    assert at == keys.size


def test_integer_term_ids_and_custom_tokenizer():
    ix = B.LexicalIndex(0)
    ix.add_token_matrix(np.array([[1, 2, 2, 9], [9, 9, 9, 9], [4, 5, 6, 7]]))
    ix.add_token_ids(np.array([2]))
    got = bm25_from_csr(ix, np.array([2, 9]))
    want = orc.bm25_scores([["1", "2", "2", "9"], ["9"] * 4, ["4", "5", "6", "7"], ["2"]], ["2", "9"])
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.allclose(got[~np.isnan(got)], want[~np.isnan(want)])
    ix2 = B.LexicalIndex(0, tokenize=lambda t: t.split("|"))                     # custom tokenizer: Python path
    assert ix2._fast is False
    ix2.add_texts(["a b|c", "c|c|d"])
    got = bm25_from_csr(ix2, "c|a b")
    want = orc.bm25_scores([["a b", "c"], ["c", "c", "d"]], ["c", "a b"])
    assert np.allclose(got, want)
    ix2.reset()
    assert len(ix2) == 0 and bm25_from_csr(ix2, "c").size == 0


def test_helper_throughput_is_reported():
    if B._text_lib() is None:
        pytest.skip("libarchi_text.so not built")
    rng = random.Random(1)
    words = ["w%d" % i for i in range(20000)]
    texts = [" ".join(rng.choice(words) for _ in range(150)) for _ in range(2000)]
    ix = B.LexicalIndex(0)
    t0 = time.perf_counter()
    ix.add_texts(texts)
    dt = time.perf_counter() - t0
    print(f"lexical add_texts (C helper): {len(texts) / dt:.0f} chunks/s")
    assert len(ix) == 2000


@pytest.mark.parametrize("fast", [True, False])
def test_score_marshals_the_postings_for_the_kernel(fast, monkeypatch):
    """LexicalIndex.score() end to end on the CPU: torch tensors live on the host and a stand-in for
    archi_bm25_accumulate applies the kernel's formula to exactly the pointers and scalars it is given."""
    import ctypes
    import types

    import torch
    from archi_b200 import _native as N
    if fast and B._text_lib() is None:
        pytest.skip("libarchi_text.so not built")
    if not fast:
        monkeypatch.setattr(B, "_text_lib", lambda: None)
    cpu = torch.device("cpu")
    monkeypatch.setattr(torch, "device", lambda *a, **k: cpu)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: types.SimpleNamespace(cuda_stream=0))

    def view(ptr, ctype, n):
        addr = ptr.value if hasattr(ptr, "value") else int(ptr)
        return np.ctypeslib.as_array((ctype * n).from_address(addr)) if n else np.empty(0)

    def fake_accumulate(starts, ends, n_terms, idf, doc_ids, tfs, doc_len, avgdl, k1, b, sign, out, stream):
        st, en = view(starts, ctypes.c_int64, n_terms), view(ends, ctypes.c_int64, n_terms)
        idf_a = view(idf, ctypes.c_float, n_terms)
        n_post, n_docs = int(ix._post_ptr[-1]), len(ix)
        d_all, tf_all = view(doc_ids, ctypes.c_int32, n_post), view(tfs, ctypes.c_int32, n_post)
        dl, o = view(doc_len, ctypes.c_float, n_docs), view(out, ctypes.c_float, n_docs)
        for t in range(n_terms):
            d, tf = d_all[st[t]:en[t]], tf_all[st[t]:en[t]].astype(np.float64)
            o[d] += (idf_a[t] * tf * (k1 + 1) / (tf + k1 * (1 - b + b * dl[d] / avgdl)) * sign).astype(np.float32)
        return 0

    monkeypatch.setattr(N, "lib", lambda: types.SimpleNamespace(archi_bm25_accumulate=fake_accumulate))
    rng = random.Random(4)
    docs = corpus(rng, 120, unicode_every=5)
    ix = B.LexicalIndex(0)
    ix.add_texts(docs)
    ix.delete_rows([7])
    live = [orc.tokenize(t) for i, t in enumerate(docs) if i != 7]
    for query in ("higgs boson quark", "the muon the", "zzz"):
        got = ix.score(query).numpy()
        want = np.nan_to_num(orc.bm25_scores(live, orc.tokenize(query)), nan=0.0)
        assert got[7] == 0.0
        assert np.allclose(np.delete(got, 7), want, rtol=1e-5, atol=1e-6), query
    out = torch.ones(len(docs))
    assert ix.score("quark", out=out) is out and out[7] == 0.0
