"""BASELINE config 1, host side: archi's docs chunked with the default data_manager settings
(manager.py:75-78,292; base-config.yaml:134-168).  CPU only."""
import os

import pytest

from archi_b200.ingest import split_text
from oracle import oracle as orc

import config1_data

REF_DOCS = "/root/reference/docs/docs"


def test_fixture_chunks_as_recorded():
    fx = config1_data.load()
    assert fx["n_files"] == 15 and fx["n_chunks"] == 120 and fx["n_chunks_over_1000_chars"] == 3
    total = 0
    for doc in fx["docs"]:
        ours = split_text(doc["text"], 1000, 0, "\n\n")
        assert [len(c) for c in ours] == doc["chunk_lengths"], doc["filename"]
        assert ours == orc.character_text_split(doc["text"], 1000, 0, "\n\n")
        total += len(ours)
    assert total == 120


@pytest.mark.skipif(not os.path.isdir(REF_DOCS), reason="/root/reference is not on this machine")
def test_real_docs_split_like_the_fixture():
    """The pseudonymised fixture cuts exactly where the reference's own docs cut."""
    fx = {d["filename"]: d for d in config1_data.load()["docs"]}
    names = sorted(f for f in os.listdir(REF_DOCS) if f.endswith(".md"))
    assert names == sorted(fx)
    for name in names:
        text = open(os.path.join(REF_DOCS, name), encoding="utf-8").read()
        ours = split_text(text, 1000, 0, "\n\n")
        assert ours == orc.character_text_split(text, 1000, 0, "\n\n")
        assert [len(c) for c in ours] == fx[name]["chunk_lengths"], name
