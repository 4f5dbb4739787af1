#!/usr/bin/env python
"""BASELINE config 1 corpus as a fixture that can travel to the GPU box.

Config 1 = archi's own docs/ (reference: docs/docs/*.md) chunked with the default data_manager
settings -- TextLoader for .md (loader_utils.py:18-37), CharacterTextSplitter("\\n\\n", 1000, 0)
(manager.py:75-78,292; base-config.yaml:134-168) -- embedded with a MiniLM-shaped encoder, top-5.

The reference tree does not exist on the GPU box and its text is not copied into this repo: every word of
the docs is replaced by a same-length pseudo-word derived from the crc32 of the lower-cased word, all
whitespace is kept byte for byte.  Chunk boundaries, chunk lengths, word repetition statistics (BM25) and
the hashing tokenizer's behaviour are therefore those of the real corpus, the text itself is gibberish.
The script asserts that the splitter cuts the pseudonymised files exactly where it cuts the real ones.

    python tests/golden/make_config1_fixture.py      # needs /root/reference
"""
import gzip
import json
import os
import re
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

DOCS = "/root/reference/docs/docs"
LETTERS = "abcdefghijklmnopqrstuvwxyz"


def pseudo_word(word: str) -> str:
    """Same length, letters only, a function of the lower-cased word."""
    h = zlib.crc32(word.lower().encode("utf-8"))
    out = []
    for i in range(len(word)):
        h = (h * 1103515245 + 12345) & 0x7FFFFFFF
        out.append(LETTERS[(h >> 8) % 26])
    return "".join(out)


def pseudonymise(text: str) -> str:
    return re.sub(r"\S+", lambda m: pseudo_word(m.group(0)), text)


def main():
    files = sorted(f for f in os.listdir(DOCS) if f.endswith(".md"))
    docs, total_chunks, long_chunks = [], 0, 0
    for name in files:
        real = open(os.path.join(DOCS, name), encoding="utf-8").read()
        fake = pseudonymise(real)
        assert len(fake) == len(real)
        real_chunks = orc.character_text_split(real, 1000, 0, "\n\n")
        fake_chunks = orc.character_text_split(fake, 1000, 0, "\n\n")
        assert [len(c) for c in real_chunks] == [len(c) for c in fake_chunks], name
        docs.append({"filename": name, "text": fake, "chunk_lengths": [len(c) for c in fake_chunks]})
        total_chunks += len(fake_chunks)
        long_chunks += sum(1 for c in fake_chunks if len(c) > 1000)
    out = os.path.join(HERE, "config1_docs.json.gz")
    payload = {"generated_by": "tests/golden/make_config1_fixture.py",
               "source": "archi docs/docs/*.md (v1.2.4), words pseudonymised, whitespace kept",
               "splitter": {"separator": "\n\n", "chunk_size": 1000, "chunk_overlap": 0},
               "n_files": len(files), "n_chunks": total_chunks, "n_chunks_over_1000_chars": long_chunks,
               "docs": docs}
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps(payload, sort_keys=True).encode("utf-8"))
    print(f"wrote {out}: {len(files)} files, {total_chunks} chunks ({long_chunks} over 1000 chars), "
          f"{os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
