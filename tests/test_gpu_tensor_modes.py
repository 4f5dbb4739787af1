"""GPU parity tests of the tensor-core path in the exact kernel modes BASELINE configs 2/3/4 run:
CTA-pair mode (cta_group::2) with D in {384, 768, 1024}, 1024 queries (8 query tiles), bf16 stores scored
as a raw cosine dot (norm_dev bound) and fp32 stores through the bf16 shadow, k = 10 (k' = 32) and
k = 100 (k' = 224), the probe launch, the threshold-tightening + resumed launch of long scans, and the
older flood-and-resume phases -- every one of them against the oracle, tie-aware and exact
(oracle.verify_topk re-scores every returned id in fp64).

The library reads its experiment knobs (ARCHI_TC_*) once per process, so the forced-phase cases run
this file as a script in a subprocess with the knobs in the environment.
Reference semantics: postgres_vectorstore.py:317-332 (ORDER BY distance ASC LIMIT k), :361 (score)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu

TENSOR = 2
REL = {"f32": 1e-5, "bf16": 2e-3}


def _rows(rng, n, d, unit):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if not unit:
        x *= rng.uniform(0.5, 1.5, size=(n, 1)).astype(np.float32)
    return x


def _stored(corpus, storage):
    return orc.bf16_bits_to_f32(orc.f32_to_bf16_bits(corpus)) if storage == "bf16" else corpus


def run_case(n, dim, storage, k, metric, unit, nq, seed=0, expect_launches=None, delete=0, path=TENSOR):
    """One corpus, one batch, tensor path, checked against the fp64 truth.  Returns the stats."""
    from archi_b200.store import NativeStore
    rng = np.random.default_rng(seed * 7919 + n + dim + nq)
    corpus = _rows(rng, n, dim, unit)
    queries = _rows(rng, nq, dim, True) * (1.0 if unit else rng.uniform(0.5, 2.0, size=(nq, 1)).astype(np.float32))
    queries = queries.astype(np.float32)
    s = NativeStore(dim, metric, storage)
    s.append(corpus)
    mask = None
    if delete:
        gone = rng.choice(n, delete, replace=False)
        s.delete_rows(gone)
        mask = np.ones(n, dtype=bool)
        mask[gone] = False
    scores, ids = s.search(queries, k, path=path)
    st = s.last_stats()
    assert st.path == path
    stored = _stored(corpus, storage)
    d_true, i_true = orc.exact_topk(metric, stored, queries, k, mask=mask, block=max(16384, (1 << 26) // nq))
    fails = orc.verify_topk(metric, stored, queries, k, ids, scores, REL[storage], d_true, i_true, mask=mask)
    assert not fails, fails[:5]
    if expect_launches is not None:
        assert st.coarse_launches == expect_launches, (st.coarse_launches, expect_launches)
    out = {"unverified": st.unverified_queries, "coarse_launches": st.coarse_launches, "grid": st.grid}
    s.close()
    return out


# (dim, storage, k, metric, unit-norm rows): the modes of configs 2 / 3 / 4 and their neighbours
PAIR_CASES = [
    (384, "f32", 10, "cosine", True),            # config 2: bf16 shadow, raw keys, split epilogue, resident query tile
    (768, "bf16", 10, "cosine", True),           # config 3: raw cosine over a bf16 store (norm_dev bound), streamed query tile
    (1024, "bf16", 100, "cosine", True),         # config 4: k' = 224
    (768, "f32", 10, "cosine", False),           # fp32 store, rows of any length (shadow rows are normalised)
    (768, "bf16", 10, "cosine", False),          # bf16 store, rows of any length: per-row constants (aux mode)
    (1024, "f32", 100, "l2", False),             # aux mode, key = 2 q.c - |c|^2
    (1024, "bf16", 100, "inner_product", False),
    (384, "bf16", 100, "l2", True),
]


@pytest.mark.parametrize("dim,storage,k,metric,unit", PAIR_CASES)
def test_pair_mode_1024_queries(dim, storage, k, metric, unit):
    """40k rows x 1024 queries: 8 query tiles -> CTA pairs, 18 groups, 157 corpus tiles >= 8 * 18 so the
    probe launch + maxima threshold run exactly as in the benchmark configurations (2 coarse launches)."""
    st = run_case(40_009, dim, storage, k, metric, unit, 1024, expect_launches=2)
    # the benchmark modes must prove (almost) every query; wide score ranges (rows of any length under l2 /
    # inner product) legitimately fail more proofs at k = 100 of 40k rows and are then answered by the exact scan
    if unit and metric == "cosine":
        assert st["unverified"] <= 10


@pytest.mark.parametrize("nq", [129, 257, 1000, 2048, 2500])
def test_pair_mode_ragged_query_counts(nq):
    """Odd tile counts are padded to even for the pair kernel; > 2048 queries take two launches."""
    run_case(20_011, 200, "bf16", 10, "cosine", True, nq)


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_probe_in_one_cta_mode(storage):
    """<= 128 queries: one CTA per SM, 148 groups; the probe needs >= 8 * 148 tiles (303k rows)."""
    run_case(310_000, 64, storage, 10, "cosine", True, 100, expect_launches=2)


def test_pair_mode_with_tombstones():
    run_case(40_009, 768, "bf16", 10, "cosine", True, 1024, delete=5000, expect_launches=2)


# ---- forced phases: the library reads ARCHI_TC_* once per process -> subprocess --------------------------
FORCED = {
    # knobs -> [(case args, expected coarse launches)]
    "tighten": ({"ARCHI_TC_P2": "2", "ARCHI_TC_PROBE": "100"},
                [((60_000, 768, "bf16", 10, "cosine", True, 1024), 3),
                 ((60_000, 1024, "bf16", 100, "cosine", True, 1024), 3),
                 ((60_000, 384, "f32", 10, "cosine", True, 1024), 3),
                 ((60_000, 256, "f32", 100, "l2", False, 1024), 3)]),
    "flood": ({"ARCHI_TC_WARM": "2"},
              [((60_000, 768, "bf16", 10, "cosine", True, 1024), 2),
               ((230_000, 64, "f32", 10, "cosine", True, 1024), 3),
               ((230_000, 64, "bf16", 100, "inner_product", False, 1024), 3)]),
    "local": ({"ARCHI_TC_WARM": "0"},
              [((60_000, 768, "bf16", 10, "cosine", True, 1024), 1),
               ((60_000, 384, "f32", 100, "cosine", True, 300), 1)]),
    "tf32": ({"ARCHI_NO_SHADOW": "1"},
             [((40_009, 384, "f32", 10, "cosine", True, 1024), 2),
              ((40_009, 100, "f32", 10, "l2", False, 200), None)]),
    "aux": ({"ARCHI_TC_RAW": "0"},
            [((40_009, 768, "bf16", 10, "cosine", True, 1024), 2),
             ((40_009, 384, "f32", 10, "inner_product", False, 1024), 2)]),
    "nopair": ({"ARCHI_TC_PAIR": "0"},
               [((40_009, 768, "bf16", 10, "cosine", True, 1024), 2)]),
    "layout": ({"ARCHI_TC_SPLIT": "0", "ARCHI_TC_ARES": "0"},
               [((40_009, 384, "f32", 10, "cosine", True, 1024), 2)]),
    "split": ({"ARCHI_TC_SPLIT": "1"},
              [((40_009, 1024, "bf16", 10, "cosine", True, 1024), 2)]),
    # soft throttle of long scans forced on for short ones: every producer publishes its tile count and waits
    # (bounded) for the slowest CTA of its corpus group -- pair mode, 1-CTA mode, tightening phase, tombstone-free
    "throttle": ({"ARCHI_TC_THROTTLE": "1"},
                 [((60_000, 768, "bf16", 10, "cosine", True, 1024), 2),
                  ((230_000, 384, "f32", 10, "cosine", True, 1024), 2),
                  ((60_000, 1024, "bf16", 100, "cosine", True, 300), None),
                  ((40_009, 384, "f32", 10, "l2", False, 100), None)]),
}


@pytest.mark.parametrize("name", sorted(FORCED))
def test_forced_phase(name):
    env = dict(os.environ)
    env.update(FORCED[name][0])
    res = subprocess.run([sys.executable, os.path.abspath(__file__), name], env=env, capture_output=True, text=True,
                         timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    print(res.stdout.strip().splitlines()[-1])


# ---- full size: one shard of config 3, one config-4-like shard, a scan long enough to tighten ---------
def _full_size(n, dim, k, nq, sample, expect_launches, seed):
    """bf16 shard generated on the device; `sample` of the nq queries are checked against oracle.c run on the
    stored bf16 bits (float accumulators, all host threads)."""
    import torch
    from archi_b200.store import NativeStore
    g = torch.Generator(device="cuda").manual_seed(seed)
    s = NativeStore(dim, "cosine", "bf16", capacity_rows=n)
    bits = np.empty((n, dim), dtype=np.uint16)
    for r0 in range(0, n, 250_000):
        m = min(250_000, n - r0)
        x = torch.randn((m, dim), generator=g, device="cuda")
        x = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
        s.append(x)
        bits[r0:r0 + m] = x.view(torch.int16).cpu().numpy().view(np.uint16)
    q = torch.randn((nq, dim), generator=g, device="cuda")
    q = q / q.norm(dim=1, keepdim=True)
    sc, ids = s.search(q, k, path=TENSOR)
    st = s.last_stats()
    torch.cuda.synchronize()
    assert st.path == TENSOR and st.unverified_queries == 0
    if expect_launches is not None:
        assert st.coarse_launches == expect_launches, st.coarse_launches
    pick = np.linspace(0, nq - 1, sample).astype(np.int64)
    qh = q.cpu().numpy()[pick]
    d_true, i_true = orc.c_scan_topk("cosine", bits, qh, k, nthreads=os.cpu_count() or 4, corpus_is_bf16=True)

    class _Rows:                      # verify_topk only gathers a few rows: up-cast them on demand
        def __getitem__(self, idx):
            return orc.bf16_bits_to_f32(bits[idx])

    fails = orc.verify_topk("cosine", _Rows(), qh, k, ids.cpu().numpy()[pick], sc.cpu().numpy()[pick], 2e-3,
                            d_true, i_true)
    assert not fails, fails[:5]
    # the streaming kernel on a few of the same queries: bit-identical ids up to ties is implied by the above;
    # here the two CUDA paths must agree on the scores they both computed in fp32
    sc1, id1 = s.search(q[:2], k, path=1)
    assert torch.allclose(sc1, sc[:2], rtol=1e-5, atol=1e-6)
    s.close()


def test_full_size_config3_shard():
    """One shard of config 3 at 8 GPUs: 1.25M x 768 bf16, 1024 queries, top-10."""
    _full_size(1_250_000, 768, 10, 1024, 64, 2, seed=3001)


def test_full_size_config4_mode():
    """Config 4's mode (D = 1024, k = 100 -> k' = 224, 1024 queries) on 1M rows."""
    _full_size(1_000_000, 1024, 100, 1024, 48, 2, seed=4001)


def test_full_size_tightening_phase():
    """>= 3.54M rows: the main scan is split once to tighten the thresholds (probe + 2 launches)."""
    _full_size(3_700_000, 128, 100, 1024, 64, 3, seed=4002)


if __name__ == "__main__":
    name = sys.argv[1]
    out = []
    for args, expect in FORCED[name][1]:
        out.append(run_case(*args, expect_launches=expect))
    print(json.dumps({"forced": name, "cases": out}))
