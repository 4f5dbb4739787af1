"""Loader of the config-1 fixture (tests/golden/config1_docs.json.gz, see make_config1_fixture.py)."""
import gzip
import json
import os

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_docs.json.gz")


def load():
    with gzip.open(FIXTURE, "rb") as f:
        return json.loads(f.read().decode("utf-8"))


def queries(chunks, n=20):
    """Deterministic queries: the first 12 words of evenly spaced chunks (a user quoting the docs)."""
    step = max(1, len(chunks) // n)
    return [" ".join(chunks[i].split()[:12]) for i in range(0, len(chunks), step)][:n]
