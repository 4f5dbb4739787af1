#!/usr/bin/env python
"""Records what the UNMODIFIED reference classes return for the conformance scenario
(tests/ref_harness.py::run_scenario) into tests/golden/reference_conformance.json.

Runs only where /root/reference exists (the build container).  The reference's
PostgresVectorStore / HybridRetriever / SemanticRetriever / GradingRetriever are executed from their own
files; the database they talk to is ref_harness.FakePg (statement shapes parsed from the SQL the reference
emits, distances and BM25 from the oracle).  Re-run after changing the scenario:

    python tests/golden/make_reference_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness as H  # noqa: E402


def main():
    impl = H.ReferenceImpl()
    transcript = H.run_scenario(impl)
    out = os.path.join(HERE, "reference_conformance.json")
    with open(out, "w") as f:
        json.dump({"generated_by": "tests/golden/make_reference_golden.py",
                   "reference": "archi-physics/archi v1.2.4: src/data_manager/vectorstore/postgres_vectorstore.py + retrievers/*.py (unmodified, loaded by path)",
                   "engines": "stand-in (ref_harness.FakePg): oracle.c distances (restated pgvector), oracle.bm25_scores (restated pg_textsearch)",
                   "transcript": transcript}, f, indent=1, sort_keys=True, ensure_ascii=False)
    print(f"wrote {out}: {len(transcript)} steps")


if __name__ == "__main__":
    main()
