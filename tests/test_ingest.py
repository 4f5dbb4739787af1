"""CPU tests of the ingestion driver (archi_b200/ingest.py): the reference's per-file semantics
(src/data_manager/vectorstore/manager.py:253-457) with cross-file, length-ordered embedding.  The GPU
store is replaced by a recording stand-in behind B200VectorStore's own bookkeeping."""
import os
import random

import numpy as np
import pytest

from archi_b200 import B200VectorStore, IngestionDriver, split_text
from oracle import oracle as orc


class FakeNative:
    """Stands in for NativeStore below B200VectorStore: keeps the appended rows on the host."""

    def __init__(self, dim):
        self.dim, self.rows, self.deleted = dim, [], []

    def append(self, emb):
        emb = np.asarray(emb, dtype=np.float32)
        assert emb.ndim == 2 and emb.shape[1] == self.dim
        first = len(self.rows)
        self.rows.extend(emb)
        return first

    def delete_rows(self, rows):
        self.deleted.extend(int(r) for r in rows)

    def close(self):
        pass


class HashEmbeddings:
    """embed_documents(text) = a vector that is a pure function of the text, so that any permutation
    mistake of the length-ordered batching shows up."""

    def __init__(self, dim=8, poison=None):
        self.dim, self.poison, self.calls = dim, poison, []

    def _one(self, t):
        rng = np.random.default_rng(abs(hash(t)) % (2 ** 32))
        return rng.standard_normal(self.dim).astype(np.float32)

    def embed_documents(self, texts):
        self.calls.append(list(texts))
        if self.poison is not None and any(self.poison in t for t in texts):
            raise RuntimeError("tokenizer exploded")
        return [self._one(t).tolist() for t in texts]

    def embed_query(self, text):
        return self._one(text).tolist()


class Catalog:
    def __init__(self):
        self.status, self.commits = {}, 0

    def update_ingestion_status(self, filehash, status, error=None):
        self.status.setdefault(filehash, []).append((status, error))

    def get_document_id(self, filehash):
        return "doc-" + filehash

    def get_metadata_for_hash(self, filehash):
        return {"url": "https://example.org/" + filehash, "none": None, 7: "seven"}

    def commit(self):
        self.commits += 1


def make_store(name, ef, monkeypatch):
    B200VectorStore.drop_collection(name)
    store = B200VectorStore({}, ef, collection_name=name, bm25_index=False)
    coll = store._coll
    fake = FakeNative(ef.dim)
    monkeypatch.setattr(coll, "ensure_native", lambda dim, shard=0: fake)
    coll.shards[0].native = fake
    coll.dim = ef.dim
    return store, fake


def write_files(tmp_path, n, rng):
    files = {}
    for i in range(n):
        paras = ["".join(rng.choice("abcdefgh ") for _ in range(rng.randint(20, 700))) for _ in range(rng.randint(1, 9))]
        p = tmp_path / f"file{i}.md"
        p.write_text("\n\n".join(paras))
        files[f"h{i}"] = str(p)
    return files


def test_split_text_matches_the_restated_splitter():
    rng = random.Random(3)
    for _ in range(200):
        paras = ["".join(rng.choice("ab \n") for _ in range(rng.randint(0, 60))) for _ in range(rng.randint(0, 12))]
        text = "\n\n".join(paras)
        for size, overlap in ((50, 0), (80, 20), (1000, 0), (30, 10)):
            assert split_text(text, size, overlap) == orc.character_text_split(text, size, overlap)
    assert split_text("") == [] and split_text("\n\n\n\n") == []
    long_piece = "x" * 1500
    assert split_text("a\n\n" + long_piece + "\n\nb", 1000, 0) == ["a", long_piece, "b"]     # oversize piece kept whole


def test_add_files_statuses_metadata_and_order(tmp_path, monkeypatch):
    rng = random.Random(5)
    files = write_files(tmp_path, 7, rng)
    (tmp_path / "blank.md").write_text(" \n\n \n\n")
    (tmp_path / "image.png").write_bytes(b"\x89PNG")
    files["hblank"], files["hpng"], files["hmissing"] = str(tmp_path / "blank.md"), str(tmp_path / "image.png"), str(tmp_path / "nope.md")
    ef, cat = HashEmbeddings(), Catalog()
    store, fake = make_store("ingest_a", ef, monkeypatch)
    report = IngestionDriver(store, catalog=cat, chunk_size=300, parallel_workers=3).add_files(files)

    assert sorted(report.embedded) == [f"h{i}" for i in range(7)]
    assert set(report.failed) == {"hblank", "hpng", "hmissing"}
    assert report.failed["hblank"] == "No text chunks could be extracted"           # manager.py:323
    assert report.failed["hpng"].startswith("Unsupported file format")              # manager.py:281
    for h in files:                                                                  # status machine, manager.py:259-261
        assert cat.status[h][0] == ("embedding", None)
        assert cat.status[h][-1][0] == ("embedded" if h in report.embedded else "failed")
    assert report.embed_calls == 1 and len(ef.calls) == 1                            # ONE cross-file embedding pass
    lengths = [len(t) for t in ef.calls[0]]
    assert lengths == sorted(lengths)                                                # length-ordered batches
    assert report.commits == 1 and cat.commits == 1

    coll = store._coll
    assert report.chunks == len(coll.texts) == len(fake.rows)
    row = 0
    for i in range(7):                                                               # rows land in file order
        chunks = [c for c in split_text(open(files[f"h{i}"]).read(), 300, 0) if c.strip()]
        for j, c in enumerate(chunks):
            assert coll.texts[row] == c
            assert np.allclose(fake.rows[row], ef._one(c))                           # every chunk got ITS embedding
            m = coll.metadatas[row]
            assert (m["chunk_index"], m["filename"], m["resource_hash"], m["collection"]) == (j, f"file{i}.md", f"h{i}", "ingest_a")
            assert m["url"] == f"https://example.org/h{i}" and m["7"] == "seven" and "none" not in m
            assert coll.document_ids[row] == f"doc-h{i}" and coll.chunk_index[row] == j
            row += 1
    B200VectorStore.drop_collection("ingest_a")


def test_one_bad_file_does_not_fail_its_group(tmp_path, monkeypatch):
    rng = random.Random(9)
    files = write_files(tmp_path, 5, rng)
    bad = tmp_path / "file2.md"
    bad.write_text(bad.read_text() + "\n\nPOISON PILL")
    ef, cat = HashEmbeddings(poison="POISON"), Catalog()
    store, fake = make_store("ingest_b", ef, monkeypatch)
    report = IngestionDriver(store, catalog=cat, chunk_size=300).add_files(files)
    assert set(report.failed) == {"h2"} and report.failed["h2"] == "tokenizer exploded"
    assert sorted(report.embedded) == ["h0", "h1", "h3", "h4"]
    assert report.group_retries == 1 and report.embed_calls == 1 + 5                # the group, then file by file
    assert cat.status["h2"][-1] == ("failed", "tokenizer exploded")
    assert all("POISON" not in t for t in store._coll.texts)                         # nothing of the bad file was stored
    B200VectorStore.drop_collection("ingest_b")


def test_commit_every_25_files_and_nul_bytes(tmp_path, monkeypatch):
    files = {}
    for i in range(60):
        p = tmp_path / f"f{i}.txt"
        p.write_text(f"para one of {i}\x00\n\npara two of {i}")
        files[f"k{i}"] = str(p)
    ef, cat = HashEmbeddings(), Catalog()
    store, fake = make_store("ingest_c", ef, monkeypatch)
    report = IngestionDriver(store, catalog=cat, chunk_size=20).add_files(files)
    assert report.commits == 3 and cat.commits == 3                                  # 25 + 25 + 10, manager.py:257,441-449
    assert report.embed_calls == 3 and len(report.embedded) == 60
    assert all("\x00" not in t for t in store._coll.texts)                           # manager.py:299-300
    assert report.chunks == 120
    assert IngestionDriver(store).add_files({}).commits == 0
    B200VectorStore.drop_collection("ingest_c")


def test_add_embedded_texts_contract(monkeypatch):
    ef = HashEmbeddings()
    store, fake = make_store("ingest_d", ef, monkeypatch)
    assert store.add_embedded_texts([], np.zeros((0, 8))) == []
    ids = store.add_embedded_texts(["a", "b"], np.ones((2, 8)), metadatas=[{"x": 1}, {}], document_id="D")
    assert len(ids) == 2 and store._coll.metadatas[0]["chunk_id"] == ids[0] and store._coll.metadatas[0]["collection"] == "ingest_d"
    with pytest.raises(ValueError, match="one vector per text"):
        store.add_embedded_texts(["a", "b"], np.ones((3, 8)))
    # upsert on (document_id, chunk_index): the old rows become tombstones (postgres_vectorstore.py:173-176)
    store.add_embedded_texts(["a2"], np.ones((1, 8)), document_id="D")
    assert fake.deleted == [0] and store._coll.live[:3] == [False, True, True]
    B200VectorStore.drop_collection("ingest_d")
