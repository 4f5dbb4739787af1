#!/usr/bin/env python
"""bench.py -- queries/sec of exact top-k retrieval on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c4s|c4] [--batch Q] [--workloads c3,c4,c5,c1,robust | none]

A "step" is one search of a Q-query batch over the whole (row-sharded) corpus: local exact top-k on
every rank, exchange + merge of the k-lists.  The headline (`value`) is BASELINE.json configs[1]
(synthetic 1M x 384 fp32 unit-norm chunks, top-10, 1024 queries), strong-scaled over N GPUs (the corpus is
fixed and row-sharded).  Prints ONE JSON line on rank 0.

  value         whole-job queries/s with queries already resident in HBM (CUDA events, max over ranks)
  e2e           the same through the public call with HOST query/result buffers (H2D + D2H timed)
  roofline      the dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  oracle.c (restated pgvector seq scan + heap top-k) on the host cores, bounded sample
  parity_checked / parity_failed   queries of the TIMED batch verified against oracle.c after timing
  workloads     sub-records for the other BASELINE configs at this N (same keys): c3 (10M x 768 bf16, top-10),
                c4 (12.5M x 1024 bf16 per GPU, top-100: configs[3] itself at N = 8), c5 (hybrid + ingest),
                c1 (archi docs, MiniLM shape, top-5), robust (clustered / duplicate-heavy corpora)
  cpu_context   the other CPU figures BASELINE.md section 3 plans (all-core BLAS search, the reference's
                per-query text serialisation, CPU encoder forward)

`--impl reference` times the CPU restatement alone (the reference's own engine -- PostgreSQL + pgvector --
cannot be installed here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows, dim, storage, k, default batch, config id, description)
    "c2": (1_000_000, 384, "f32", 10, 1024, 2, "configs[1]: synthetic 1M x 384 fp32 unit-norm chunks, top-10"),
    "c2s8": (125_000, 384, "f32", 10, 1024, 2, "one shard of configs[1] at 8 GPUs (125k x 384 fp32) on one GPU: exercises the small-shard timing policy (rotating shard copies), not a bench line"),
    "c3": (10_000_000, 768, "bf16", 10, 1024, 3, "configs[2]: synthetic 10M x 768 bf16 chunks, top-10, row-sharded"),
    "c4s": (12_500_000, 1024, "bf16", 100, 1024, 4, "configs[3] one shard: 12.5M x 1024 bf16 chunks per GPU (of 100M over 8), top-100"),
    "c4": (100_000_000, 1024, "bf16", 100, 1024, 4, "configs[3]: synthetic 100M x 1024 bf16 chunks (204.8 GB) row-sharded, top-100"),
}
METRIC = "queries_per_sec_exact_top10"   # top-100 for the c4 workloads (config.k says which)
L2_BYTES = 126e6


def profiled_traffic(kernel, workload, batch):
    """dram bytes per launch of the dominant kernel from the committed ncu captures, or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            ent = json.load(open(os.path.join(ROOT, "profiles", name))).get(f"{kernel}|{workload}|{batch}")
            if ent:
                return ent["traffic_bytes"]
        except Exception:
            continue
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# synthetic corpora (device-generated; --data selects the distribution)
# ---------------------------------------------------------------------------------------------------------
def gen_rows_device(n, d, seed, device, data="gaussian", chunk=262144):
    """Unit-norm fp32 rows in chunks.  gaussian: isotropic.  latent: a 12-d latent mixed into d
    dimensions + 5 % noise (embedding-like: clustered neighbours, small score gaps).  dupes: 1/16 of the rows are
    distinct, the others are copies of them with 1e-3 noise (near-duplicate-heavy corpus)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    ga = torch.Generator(device=device).manual_seed(99)        # the mixing matrix is shared by rows and queries
    A = torch.randn((12, d), generator=ga, device=device) if data == "latent" else None
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        if data == "latent":
            x = torch.randn((m, 12), generator=g, device=device) @ A + 0.05 * torch.randn((m, d), generator=g, device=device)
        elif data == "dupes":
            base = torch.randn((max(1, m // 16), d), generator=g, device=device)
            x = base.repeat_interleave(16, dim=0)[:m] if base.shape[0] * 16 >= m else base.repeat((m + base.shape[0] - 1) // base.shape[0], 1)[:m]
            x = x / x.norm(dim=1, keepdim=True) + 1e-3 * torch.randn((m, d), generator=g, device=device)
        else:
            x = torch.randn((m, d), generator=g, device=device, dtype=torch.float32)
        yield x / x.norm(dim=1, keepdim=True)


def gen_queries_device(q, d, seed, device, data="gaussian"):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    if data == "latent":
        ga = torch.Generator(device=device).manual_seed(99)
        A = torch.randn((12, d), generator=ga, device=device)
        x = torch.randn((q, 12), generator=g, device=device) @ A + 0.05 * torch.randn((q, d), generator=g, device=device)
    else:
        x = torch.randn((q, d), generator=g, device=device, dtype=torch.float32)
    return x / x.norm(dim=1, keepdim=True)


# ---------------------------------------------------------------------------------------------------------
# CPU legs
# ---------------------------------------------------------------------------------------------------------
def host_corpus(rows, dim, cfg_id):
    rng = np.random.default_rng(1234 + 1000 * cfg_id)
    corpus = np.empty((rows, dim), dtype=np.float32)
    for s in range(0, rows, 131072):
        e = min(rows, s + 131072)
        x = rng.standard_normal((e - s, dim), dtype=np.float32)
        corpus[s:e] = x / np.linalg.norm(x, axis=1, keepdims=True)
    return corpus, rng


def cpu_reference_leg(rows, dim, k, target_s, cfg_id):
    """oracle.c on all host threads over a bounded sample of the workload: the full corpus (fp32 on
    the host, as the reference stores float4), `sample_q` queries of the batch."""
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    corpus, rng = host_corpus(rows, dim, cfg_id)
    q = rng.standard_normal((max(threads * 64, 64), dim), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t0 = time.perf_counter()
    orc.c_scan_topk("cosine", corpus, q[:threads], k, nthreads=threads, fast=True)      # calibration = warm-up
    per_round = time.perf_counter() - t0
    rounds = int(max(1, min(64, target_s / max(per_round, 1e-3))))
    sample_q = min(q.shape[0], threads * rounds)
    return orc, corpus, q, threads, sample_q


REF_NOTE = ("oracle.c seq scan + heap top-k built with pgvector's flags (-O3 -march=native -fassociative-math), one query "
            "per host thread; the reference engine (PostgreSQL + pgvector) is not installable here")


def base_config(desc, rows, dim, k, batch, storage):
    """The workload, identically keyed in both arms (everything about HOW it ran goes under `run`)."""
    return {"workload": desc, "rows": rows, "dim": dim, "k": k, "batch": batch, "storage": storage, "metric": "cosine"}


def run_reference(args, rows, dim, storage, k, batch, desc, cfg_id):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc, corpus, q, threads, sample_q = cpu_reference_leg(rows, dim, k, target_s=8.0, cfg_id=cfg_id)
    for _ in range(args.warmup):
        orc.c_scan_topk("cosine", corpus, q[:threads], k, nthreads=threads, fast=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.c_scan_topk("cosine", corpus, q[:sample_q], k, nthreads=threads, fast=True)
    dt = time.perf_counter() - t0
    qps = sample_q * args.steps / dt
    sample = (f"each timed step scans the whole {rows}x{dim} fp32 corpus for a sample of {sample_q} queries "
              f"(a full step is {batch}); ms_per_step is scaled to the {batch}-query batch; {REF_NOTE}")
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": batch / qps * 1e3,
            "sample_ms": dt / args.steps * 1e3, "sample_queries": sample_q,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(desc, rows, dim, k, batch, storage),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json_line(line)


def cpu_context(rows, dim, k, batch, cfg_id):
    """the other CPU figures next to the GPU numbers (rank 0, N = 1 only)."""
    import torch
    out = {}
    threads = os.cpu_count() or 1
    # (1) cpu_blas: all-core fp32 corpus @ q^T + top-k -- the best case for an exact search on the host
    try:
        corpus, rng = host_corpus(rows, dim, cfg_id)
        q = rng.standard_normal((batch, dim), dtype=np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        c_t, q_t = torch.from_numpy(corpus), torch.from_numpy(q)
        torch.set_num_threads(threads)
        qs = min(batch, 256)

        def one():
            best_v, best_i = None, None
            for s in range(0, rows, 262144):
                sc = q_t[:qs] @ c_t[s:s + 262144].T
                v, i = torch.topk(sc, k, dim=1)
                i = i + s
                if best_v is None:
                    best_v, best_i = v, i
                else:
                    v2, j = torch.topk(torch.cat([best_v, v], 1), k, dim=1)
                    best_v, best_i = v2, torch.gather(torch.cat([best_i, i], 1), 1, j)
            return best_v, best_i
        one()
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 4.0:
            one()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        out["cpu_blas"] = {"value": qs / dt, "unit": "queries/s", "cores": threads,
                           "sample": f"{qs} of {batch} queries x {rows}x{dim} fp32, torch.matmul (MKL/oneDNN sgemm) + torch.topk, blocked over rows"}
        del corpus, c_t
    except Exception as e:  # noqa: BLE001
        out["cpu_blas"] = {"error": str(e)}
    # (2) cpu_ref_overhead: the reference serialises every query vector as decimal text (postgres_vectorstore.py:313,391)
    over = {}
    for d in (384, 768, 1024):
        v = np.random.default_rng(d).standard_normal(d).astype(np.float32).tolist()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 0.3:
            s = "[" + ",".join(str(x) for x in v) + "]"
            n += 1
        over[str(d)] = {"us_per_query": (time.perf_counter() - t0) / n * 1e6, "text_bytes": len(s)}
    out["cpu_ref_overhead"] = {"what": "'[' + ','.join(str(x) for x in embedding) + ']' per query, 1 core", "by_dim": over}
    return out


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON record: keep the real stdout aside and point fd 1 at
    stderr for everything else (NCCL prints its version banner on stdout from native code)."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(line) -> None:
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


# ---------------------------------------------------------------------------------------------------------
# one dense workload on the current process group
# ---------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, world, rank, local_rank, dev, peaks):
        self.world, self.rank, self.local_rank, self.dev, self.peaks = world, rank, local_rank, dev, peaks

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())


def run_dense(env, name, total_rows, dim, storage, k, batch, cfg_id, desc, steps, warmup, *, data="gaussian",
              sub_batches=(), parity_n=16, with_e2e=True, roofline_workload=None):
    """Builds the row-sharded store for one workload, times it, verifies sampled queries of the timed batch
    against oracle.c, and returns the record (rank 0) -- None on the other ranks."""
    import torch
    import torch.distributed as dist
    from archi_b200 import _native as N
    from archi_b200.sharded import ShardedStore, plan_row_shards
    from archi_b200.store import NativeStore
    world, rank, dev, peaks = env.world, env.rank, env.dev, env.peaks
    first, cnt = plan_row_shards(total_rows, world)[rank]
    store = NativeStore(dim, "cosine", storage, device=env.local_rank, capacity_rows=cnt)
    # the host keeps the stored values of this rank's shard for the parity check (bf16 bits for bf16 stores)
    keep_host = parity_n > 0
    parity_note = None
    if keep_host:
        # every rank of the box keeps its shard on the host for the check: bounded by the RAM that is actually free
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 1 << 62
        need = cnt * dim * (2 if storage == "bf16" else 4)
        fits = 0.0 if need * max(world, 1) > 0.6 * avail else 1.0
        if world > 1:                       # one decision for the whole job: the check holds collectives
            t_fit = torch.tensor([fits], dtype=torch.float64, device=dev)
            dist.all_reduce(t_fit, op=dist.ReduceOp.MIN)
            fits = float(t_fit.item())
        if fits < 0.5:
            keep_host, parity_n = False, 0
            parity_note = (f"skipped: the shards' host copies ({need * world / 1e9:.0f} GB over the box's ranks) do not fit "
                           f"in the free host memory ({avail / 1e9:.0f} GB)")
    host = np.empty((cnt, dim), dtype=np.uint16 if storage == "bf16" else np.float32) if keep_host else None
    at = 0
    for x in gen_rows_device(cnt, dim, 1234 + 1000 * cfg_id + rank, dev, data):
        store.append(x)
        if keep_host:
            if storage == "bf16":
                host[at:at + x.shape[0]] = x.to(torch.bfloat16).view(torch.int16).cpu().numpy().view(np.uint16)
            else:
                host[at:at + x.shape[0]] = x.cpu().numpy()
        at += x.shape[0]
    sharded = ShardedStore(store)
    sharded.sync_layout(device=dev)
    assert sharded.total_rows == total_rows
    elt = 2 if storage == "bf16" else 4

    # Timing rule: inputs larger than L2, or an L2 flush between timed iterations.  What a step reads per GPU is the
    # shard (the bf16 shadow of an fp32 shard on the tensor path).  When that is not at least 2x the 126 MB L2 (strong
    # scaling shrinks it), the shard is held R times at different addresses (identical rows, so identical answers) and
    # consecutive steps scan consecutive copies: R x shard bytes >= 2x L2 pass through the cache between two reads of
    # the same bytes -- "inputs larger than L2" without a flush + synchronize after every step (which would expose the
    # host's launch latency instead of the GPU's time and needed a flush-only loop subtracted).
    # Decided from the largest shard so that every rank takes the same branch.
    shard_read_bytes = (-(-total_rows // world)) * dim * (2 if storage == "bf16" or os.environ.get("ARCHI_NO_SHADOW", "0") == "0" else 4)
    n_copies = 1
    if shard_read_bytes < 2 * L2_BYTES:
        n_copies = int(-(-2 * L2_BYTES // max(shard_read_bytes, 1)))
        n_copies = max(2, min(n_copies, 16))
    copies = [store]
    for _ in range(n_copies - 1):
        extra = NativeStore(dim, "cosine", storage, device=env.local_rank, capacity_rows=cnt)
        for x in gen_rows_device(cnt, dim, 1234 + 1000 * cfg_id + rank, dev, data):
            extra.append(x)
        copies.append(extra)
    turn = [0]
    host_ms = [0.0]

    def next_copy():
        """The shard copy this step scans (round robin); the exchange and the id offsets are shared."""
        c = copies[turn[0] % n_copies]
        turn[0] += 1
        sharded.native = c
        return c

    def time_device(q_dev, n_steps, n_warm):
        for _ in range(max(n_warm, n_copies)):
            next_copy()
            sharded.search(q_dev, k)
        env.barrier()
        if world > 1:
            # The ranks leave the barrier up to a few milliseconds apart on the host; the first exchange of the timed
            # region would charge that skew to the early ranks (it is a property of the barrier, not of a step, and
            # with K = 20 steps it is the larger part of the measurement).  Two more untimed steps couple the GPUs'
            # streams through their exchanges before the first event is recorded.
            for _ in range(2):
                next_copy()
                sharded.search(q_dev, k)
        l0 = N.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for _ in range(n_steps):
            next_copy()
            sharded.search(q_dev, k)
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / n_steps      # host time to enqueue a step (no sync inside)
        e1.record()
        env.barrier()
        launches = N.kernel_launches() - l0
        total = e0.elapsed_time(e1)
        return env.max_over_ranks(total) / n_steps, launches

    def kernel_roofline(q_dev, reps=5):
        """Average CUDA-event duration of the dominant kernel (events recorded inside the library
        around that launch, on the stream it is launched on) and its algorithmic bytes."""
        ms, st = [], None
        for _ in range(max(reps, n_copies)):
            c = next_copy()
            c.set_timing(True)
            c.search(q_dev, k)
            st = c.last_stats()
            ms.append(st.last_kernel_ms)
            c.set_timing(False)
        launch_ms = float(np.median(ms))
        nq_all = q_dev.shape[0]
        wl = roofline_workload or name
        if st.path == N.PATH_STREAM:
            nq_launch = min(nq_all, 8)
            algo_bytes = cnt * dim * elt + nq_launch * dim * 4 + nq_launch * k * 8
            ach = algo_bytes / (launch_ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"],
                    "traffic": profiled_traffic("scan_topk_kernel", wl, nq_all) if world == 1 else None,
                    "peak_source": peaks["source"] + " hbm_gbs (copy, burst)", "kernel": "scan_topk_kernel",
                    "launch_ms": launch_ms, "algorithmic_bytes_per_launch": algo_bytes,
                    "launches_per_step": st.passes, "grid": st.grid, "frac_of_nominal_8TBps": ach / 8000.0,
                    "unverified_queries": 0}
        # tensor-core path: one pass scores min(nq, 2048) queries against the whole shard; the coarse
        # kernel reads bf16 rows (bf16 stores, or the bf16 shadow of an fp32 store) or fp32 rows as tf32
        nq_launch = min(nq_all, 2048)
        celt = 2 if st.coarse_dtype == N.BF16 else 4
        algo_bytes = cnt * dim * celt + nq_launch * dim * celt
        algo_flops = 2.0 * nq_launch * cnt * dim
        gbs = algo_bytes / (launch_ms * 1e-3) / 1e9
        tfs = algo_flops / (launch_ms * 1e-3) / 1e12
        t_hbm = algo_bytes / (peaks["hbm_gbs"] * 1e9)
        t_tc = algo_flops / (peaks["bf16_tflops"] * 1e12)
        kname = "tc_coarse_kernel<bf16>" if st.coarse_dtype == N.BF16 else "tc_coarse_kernel<tf32>"
        if st.coarse_launches > 1:
            kname += (f" x{st.coarse_launches} (probe launch over ~1/12 of the rows + main scan) + threshold kernel "
                      f"x{st.coarse_launches - 1}, timed together; the probe's flops/bytes are NOT counted as algorithmic work")
        common = {"traffic": profiled_traffic("tc_coarse_kernel", wl, nq_all) if world == 1 else None,
                  "kernel": kname, "coarse_reads": "bf16 shadow of the fp32 rows" if (storage == "f32" and st.coarse_dtype == N.BF16) else storage + " rows",
                  "launch_ms": launch_ms, "algorithmic_bytes_per_launch": algo_bytes,
                  "algorithmic_flops_per_launch": algo_flops, "launches_per_step": st.passes, "grid": st.grid,
                  "achieved_GBps": gbs, "achieved_TFLOPs": tfs, "frac_hbm": gbs / peaks["hbm_gbs"],
                  "frac_tensor_bf16_peak": tfs / peaks["bf16_tflops"],
                  "frac_tensor_bf16_sustained_peak": tfs / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]),
                  "unverified_queries": st.unverified_queries, "unproven_queries": st.unproven_queries}
        if t_tc >= t_hbm:
            note = " (kind::tf32 runs at half the bf16 rate; frac is against the bf16 peak)" if st.coarse_dtype != N.BF16 else ""
            return dict(common, bound="tensor", achieved=tfs, peak=peaks["bf16_tflops"], unit="TFLOP/s",
                        frac=tfs / peaks["bf16_tflops"],
                        peak_source=peaks["source"] + " bf16_tflops (cuBLAS 8192^3, burst)" + note)
        return dict(common, bound="hbm", achieved=gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=gbs / peaks["hbm_gbs"],
                    peak_source=peaks["source"] + " hbm_gbs (copy, burst)", frac_of_nominal_8TBps=gbs / 8000.0)

    def time_e2e(q_host, n_steps, n_warm):
        """Public call with host buffers: pinned queries -> H2D -> search -> exchange/merge -> D2H."""
        q_pin = torch.from_numpy(q_host).pin_memory()
        out_s = torch.empty((q_host.shape[0], k), dtype=torch.float32).pin_memory()
        out_i = torch.empty((q_host.shape[0], k), dtype=torch.int64).pin_memory()
        if world == 1:
            # the C-ABI call itself takes the host buffers (archi_search with ARCHI_HOST in/out)
            q_np, outs = q_pin.numpy(), (out_s.numpy(), out_i.numpy())

            def one():
                next_copy().search(q_np, k, out=outs)
        else:
            def one():
                next_copy()
                q_d = q_pin.to(dev, non_blocking=True)
                s, i = sharded.search(q_d, k)
                out_s.copy_(s, non_blocking=True)
                out_i.copy_(i, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        for _ in range(max(n_warm, n_copies)):
            one()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            one()
        env.barrier()
        total = (time.perf_counter() - t0) * 1e3
        return env.max_over_ranks(total) / n_steps

    def parity_check(q_dev, n_check):
        """Sampled queries of the timed batch against oracle.c: every rank scans ITS shard's stored values on
        the host (k + 8 best, float accumulators), rank 0 merges the shards' lists into the global truth and
        compares the GPU's merged answer with it, tie-aware (an id outside the truth's top-k must sit within the
        tie tolerance of the k-th distance in the extended truth list).  Returns (checked, failed)."""
        from oracle import oracle as orc
        nq_all = q_dev.shape[0]
        pick = np.unique(np.linspace(0, nq_all - 1, min(n_check, nq_all)).astype(np.int64))
        sc, ids = sharded.search(q_dev, k)
        torch.cuda.synchronize()
        qh = q_dev.cpu().numpy()[pick]
        kk = k + 8
        if cnt > 0:
            d_loc, i_loc = orc.c_scan_topk("cosine", host, qh, min(kk, cnt), nthreads=max(1, (os.cpu_count() or 4) // max(1, min(world, 8))),
                                           corpus_is_bf16=(storage == "bf16"))
            pad = kk - d_loc.shape[1]
            if pad > 0:
                d_loc = np.concatenate([d_loc, np.full((len(pick), pad), np.inf)], 1)
                i_loc = np.concatenate([i_loc, np.full((len(pick), pad), -1, dtype=np.int64)], 1)
            i_loc = np.where(i_loc >= 0, i_loc + first, -1)
        else:
            d_loc, i_loc = np.full((len(pick), kk), np.inf), np.full((len(pick), kk), -1, dtype=np.int64)
        if world > 1:
            td = torch.from_numpy(d_loc).to(dev)
            ti = torch.from_numpy(i_loc).to(dev)
            gd = [torch.empty_like(td) for _ in range(world)]
            gi = [torch.empty_like(ti) for _ in range(world)]
            dist.all_gather(gd, td)
            dist.all_gather(gi, ti)
            d_all = torch.cat(gd, 1).cpu().numpy()
            i_all = torch.cat(gi, 1).cpu().numpy()
        else:
            d_all, i_all = d_loc, i_loc
        if rank != 0:
            return len(pick), 0
        order = np.lexsort((i_all, d_all), axis=1)               # distance ascending, then id
        d_all, i_all = np.take_along_axis(d_all, order, 1), np.take_along_axis(i_all, order, 1)
        got_i, got_s = ids.cpu().numpy()[pick], sc.cpu().numpy()[pick]
        rel = 2e-3 if storage == "bf16" else 1e-5
        failed = 0
        for r in range(len(pick)):
            truth_i, truth_d = i_all[r, :k], d_all[r, :k]
            kth = truth_d[-1]
            tol = 1e-7 + 2e-6 * abs(kth)
            ext = {int(i): float(d) for i, d in zip(i_all[r], d_all[r])}
            got = [int(x) for x in got_i[r]]
            ok = len(set(got)) == k and {int(i) for i, d in zip(truth_i, truth_d) if d < kth - tol}.issubset(got)
            ok = ok and all(g in ext and ext[g] <= kth + tol for g in got)
            ok = ok and np.allclose(got_s[r], 1.0 - truth_d, rtol=rel, atol=rel * 0.1)
            failed += 0 if ok else 1
        return len(pick), failed

    q_dev = gen_queries_device(batch, dim, 4321 + 1000 * cfg_id, dev, data)
    clk = ClockSampler(env.local_rank)
    clk.__enter__()          # sampled across every timed region of this workload
    ms_step, launches = time_device(q_dev, steps, warmup)
    host_enqueue_ms = host_ms[0]
    roof = kernel_roofline(q_dev) if cnt > 0 else None
    ms_e2e = time_e2e(q_dev.cpu().numpy(), steps, warmup) if with_e2e else None
    batches = {}
    for qb in sub_batches:
        if qb == batch:
            continue
        qd = gen_queries_device(qb, dim, 4321 + 1000 * cfg_id, dev, data)
        steps_b = max(steps, 50 if qb <= 8 else steps)
        ms_b, _ = time_device(qd, steps_b, warmup)
        rec = {"queries_per_s": qb / (ms_b * 1e-3), "ms_per_step": ms_b, "roofline": kernel_roofline(qd)}
        if with_e2e:
            rec["e2e_queries_per_s"] = qb / (time_e2e(qd.cpu().numpy(), steps_b, warmup) * 1e-3)
        if parity_n > 0:
            rec["parity_checked"], rec["parity_failed"] = parity_check(qd, min(parity_n, qb))
        batches[str(qb)] = rec
    clk.__exit__()
    clocks = clk.summary()
    checked, failed = parity_check(q_dev, parity_n) if parity_n > 0 else (0, 0)
    sharded.check()      # a peer-memory exchange that timed out would have produced invalid results
    exchange_kind = sharded.exchange_kind
    sharded.close()
    for c in copies:
        c.close()
    del host
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    rec = {
        "metric": METRIC if k == 10 else f"queries_per_sec_exact_top{k}", "value": batch / (ms_step * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "dtype": ("f32 (bf16 tensor-core candidate pass over a bf16 shadow, exact fp32 rescoring of every returned row)"
                  if storage == "f32" else "bf16 storage, fp32 accumulate (exact fp32 rescoring of every returned row)"),
        "data": "synthetic" if data == "gaussian" else f"synthetic ({data})",
        "config": base_config(desc, total_rows, dim, k, batch, storage),
        "run": {"rows_per_gpu": cnt, "sharding": f"rows/{world}",
                "shard_exchange": ("none (1 GPU)" if world == 1 else
                                   "one kernel: push k-lists into peers' HBM over NVLink (CUDA IPC), wait, merge"
                                   if exchange_kind == "peer-memory" else
                                   "NCCL all_gather_into_tensor of packed k-lists + merge kernel"),
                "l2_policy": (f"a step reads {shard_read_bytes / 1e6:.0f} MB per GPU vs 126 MB L2: no flush needed"
                              if n_copies == 1 else
                              f"a step reads only {shard_read_bytes / 1e6:.0f} MB per GPU: consecutive steps scan {n_copies} "
                              f"copies of the shard held at different addresses in turn ({n_copies * shard_read_bytes / 1e6:.0f} MB "
                              "pass through the 126 MB L2 between two reads of the same bytes); no flush, no per-step synchronize")},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms, "roofline": roof, "clocks": clocks,
        "unverified_queries": (roof or {}).get("unverified_queries", 0),
        "parity_checked": checked, "parity_failed": failed,
        "parity_how": "sampled queries of the timed batch vs oracle.c over the stored values of every shard (host), merged on rank 0, tie-aware",
    }
    if parity_note:
        rec["parity_how"] = parity_note
    if with_e2e:
        rec["e2e"] = {"value": batch / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": batch * dim * 4,
                      "d2h_bytes_per_step": batch * k * 12, "ms_per_step": ms_e2e}
    if batches:
        rec["batches"] = batches
    return rec


# ---------------------------------------------------------------------------------------------------------
# config 5 (hybrid + ingest) and config 1 (archi docs) -- rank 0's GPU only
# ---------------------------------------------------------------------------------------------------------
def run_c5(env, rows=1_000_000, dim=384, n_queries=100):
    """BASELINE configs[4]: hybrid BM25 + dense on 1M synthetic chunks (k = 5, weights 0.4 / 0.6) and end-to-end
    ingest (encoder forward + fused pool/normalise/append) in chunks/s.  Documents are 24 Zipf(1.3) term ids over a
    50k vocabulary; two query classes: 'content' terms (rank >= 100: what survives a stop-word list, a few % of the
    rows match) and 'any' terms (the head of the Zipf law: most rows match)."""
    import torch
    from archi_b200.bm25 import LexicalIndex
    from archi_b200.store import NativeStore, pool_normalize
    from oracle import oracle as orc
    dev, hbm_peak = env.dev, env.peaks["hbm_gbs"]
    out = {"config": {"workload": "configs[4]: hybrid BM25 + dense on synthetic chunks, plus ingest", "rows": rows, "dim": dim,
                      "k": 5, "batch": 1, "storage": "f32", "metric": "cosine", "semantic_weight": 0.4, "bm25_weight": 0.6}}
    store = NativeStore(dim, "cosine", "f32", device=env.local_rank, capacity_rows=rows)
    host = np.empty((rows, dim), dtype=np.float32)
    at = 0
    for x in gen_rows_device(rows, dim, 1234 + 5000, dev):
        store.append(x)
        host[at:at + x.shape[0]] = x.cpu().numpy()
        at += x.shape[0]
    rng = np.random.default_rng(5)
    vocab, doc_len = 50_000, 24
    t0 = time.perf_counter()
    lex = LexicalIndex(device=env.local_rank)
    tokens = (rng.zipf(1.3, size=(rows, doc_len)) % vocab).astype(np.int64)
    lex.add_token_matrix(tokens)
    classes = {"content_terms": [((rng.zipf(1.3, size=3) % (vocab - 100)) + 100).astype(np.int64) for _ in range(n_queries)],
               "any_terms": [(rng.zipf(1.3, size=3) % vocab).astype(np.int64) for _ in range(n_queries)]}
    lex.score(classes["any_terms"][0])                           # builds the device posting lists
    torch.cuda.synchronize()
    out["bm25_index_build_s"] = time.perf_counter() - t0
    q = gen_queries_device(n_queries, dim, 4321 + 5000, dev)

    def timed(fn):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_queries):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n_queries

    ms_dense = timed(lambda i: store.search(q[i:i + 1], 5))
    out["dense_only"] = {"queries_per_s": 1e3 / ms_dense, "ms_per_query": ms_dense}
    qh = q.cpu().numpy()
    for cname, terms in classes.items():
        ms = timed(lambda i: store.hybrid_search_terms(lex, [terms[i]], q[i:i + 1], 5, 0.4, 0.6))
        rec = {"queries_per_s": 1e3 / ms, "ms_per_query": ms, "vs_dense_only": ms_dense / ms}
        # batch of 64 queries in one call
        nb = min(64, n_queries)
        for _ in range(2):
            store.hybrid_search_terms(lex, terms[:nb], q[:nb], 5, 0.4, 0.6)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            store.hybrid_search_terms(lex, terms[:nb], q[:nb], 5, 0.4, 0.6)
        e1.record()
        torch.cuda.synchronize()
        rec["batch64_queries_per_s"] = nb * 5 / (e0.elapsed_time(e1) * 1e-3)
        # parity: 8 queries against the oracle's hybrid over the same stored rows and the restated BM25
        checked = failed = 0
        matched, paths = [], {}
        for i in range(0, n_queries, max(1, n_queries // 8)):
            sc, ids = store.hybrid_search_terms(lex, [terms[i]], q[i:i + 1], 5, 0.4, 0.6)
            paths[store.last_hybrid_path] = paths.get(store.last_hybrid_path, 0) + 1
            bm = lex.score(terms[i]).cpu().numpy().astype(np.float64)
            matched.append(int((bm != 0).sum()))
            comb, cid = orc.c_hybrid_topk("cosine", host, qh[i], np.where(bm != 0, bm, np.nan), 0.4, 0.6, 5)
            got_i, got_s = ids[0].cpu().numpy(), sc[0].cpu().numpy()
            ok = sorted(got_i.tolist()) == sorted(cid.tolist()) and np.allclose(np.sort(got_s), np.sort(comb), rtol=1e-5, atol=2e-6)
            checked += 1
            failed += 0 if ok else 1
        rec.update(parity_checked=checked, parity_failed=failed, rows_with_bm25_match_mean=float(np.mean(matched)),
                   paths_of_sampled_queries=paths)
        out["hybrid_" + cname] = rec
    store.close()
    del host, tokens
    # ---------------- ingest ----------------
    from transformers import BertConfig, BertModel
    from archi_b200.embeddings import MINILM_L6
    torch.manual_seed(0)
    model = BertModel(BertConfig(**MINILM_L6), add_pooling_layer=False).to(dev, torch.bfloat16).eval()
    B, L, H = 1024, 256, MINILM_L6["hidden_size"]
    ids = torch.randint(1000, 30000, (B, L), device=dev)
    lens = torch.randint(L // 2, L + 1, (B,), device=dev)
    mask = (torch.arange(L, device=dev)[None, :] < lens[:, None]).to(torch.int64)
    n_batches = 8
    sink = NativeStore(H, "cosine", "bf16", device=env.local_rank, capacity_rows=B * (n_batches + 4))

    def step():
        with torch.inference_mode():
            hidden = model(input_ids=ids, attention_mask=mask).last_hidden_state
        sink.pool_normalize_append(hidden, mask)
        return hidden

    for _ in range(2):
        hidden = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_batches):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / n_batches
    # the pool kernel alone: a CUDA graph of 20 launches, so that no host time sits between them (a launch is
    # shorter than the Python call that issues it)
    def time_pool(h_, m_):
        for _ in range(3):
            pool_normalize(h_, m_, want_bf16=True)
        torch.cuda.synchronize()
        how = "cuda graph of 20 launches"
        try:
            side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                pool_normalize(h_, m_, want_bf16=True)       # first use of this stream (lazy per-stream scratch)
                side.synchronize()
                with torch.cuda.graph(graph, stream=side):
                    for _ in range(20):
                        pool_normalize(h_, m_, want_bf16=True)
            best = 1e9
            for _ in range(5):
                torch.cuda.synchronize()
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 20)
            del graph
        except Exception as exc:                                   # capture refused: time the plain loop (host-bound)
            how = f"loop of 20 Python calls (graph capture failed: {type(exc).__name__})"
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                pool_normalize(h_, m_, want_bf16=True)
            e1.record()
            torch.cuda.synchronize()
            best = e0.elapsed_time(e1) / 20
        return best, how

    ms_pool, pool_timing = time_pool(hidden, mask)
    # the same kernel on the batch a 1M-token embedding budget produces (4096 sequences)
    B4 = 4096
    hidden4 = torch.randn((B4, L, H), device=dev, dtype=torch.float32).to(torch.bfloat16)
    lens4 = torch.randint(L // 2, L + 1, (B4,), device=dev)
    mask4 = (torch.arange(L, device=dev)[None, :] < lens4[:, None]).to(torch.int64)
    ms_pool4, _ = time_pool(hidden4, mask4)
    bytes4 = int(mask4.sum().item()) * H * 2 + B4 * L * 8 + B4 * H * (4 + 2)
    del hidden4, mask4
    live_tokens = int(mask.sum().item())
    pool_bytes = live_tokens * H * 2 + B * L * 8 + B * H * (4 + 2)   # masked tokens are not read
    out["ingest"] = {"chunks_per_s": B / (ms_step * 1e-3), "ms_per_batch": ms_step, "batch": B, "seq_len": L,
                     "encoder": "BertModel MiniLM-L6 shape, random init, bf16 (PyTorch)",
                     "pool_normalize_ms": ms_pool, "pool_timing": pool_timing, "pool_algorithmic_bytes": pool_bytes,
                     "pool_achieved_GBps": pool_bytes / (ms_pool * 1e-3) / 1e9,
                     "pool_frac_hbm": pool_bytes / (ms_pool * 1e-3) / 1e9 / hbm_peak,
                     "pool_share_of_step": ms_pool / ms_step,
                     "pool_batch4096": {"ms": ms_pool4, "algorithmic_bytes": bytes4,
                                        "achieved_GBps": bytes4 / (ms_pool4 * 1e-3) / 1e9,
                                        "frac_hbm": bytes4 / (ms_pool4 * 1e-3) / 1e9 / hbm_peak}}
    sink.close()
    del model
    torch.cuda.empty_cache()
    # the ingestion DRIVER end to end: synthetic markdown files -> load + split -> cross-file,
    # length-ordered, token-budget embedding -> fused pool/append + lexical index -> statuses and commits; beside it
    # the reference's loop shape (one embedding call and one commit per file, manager.py:362-373)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_ingest_driver", os.path.join(ROOT, "tools", "bench_ingest_driver.py"))
        bid = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bid)
        from archi_b200 import B200Embeddings
        ef = B200Embeddings(device=env.local_rank)
        grouped = bid.run(n_files=200, ef=ef, device=env.local_rank)
        per_file = bid.run(n_files=200, per_file=True, ef=ef, device=env.local_rank)
        out["ingest_driver"] = {"chunks_per_s": grouped["value"], "per_file_loop_chunks_per_s": per_file["value"],
                                "files": grouped["files"], "chunks": grouped["chunks"], "embed_calls": grouped["embed_calls"],
                                "per_file_embed_calls": per_file["embed_calls"], "failed": grouped["failed"] + per_file["failed"],
                                "what": "IngestionDriver.add_files on synthetic .md files (hash tokenizer, MiniLM-L6-shape encoder, "
                                        "bf16 store, BM25 index on): wall clock incl. file reads, splitting, tokenising"}
        del ef
        torch.cuda.empty_cache()
    except Exception as exc:  # noqa: BLE001
        out["ingest_driver"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def run_c1(env):
    """BASELINE configs[0]: archi's own docs (pseudonymised fixture with identical chunking, tests/golden/
    config1_docs.json.gz) -> split_text -> MiniLM-shaped encoder -> fused pool/normalise/append -> top-5 through the
    B200VectorStore surface; beside it the same path restated on the CPU (torch fp32 encoder = cpu_embed, oracle.c)."""
    import copy
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import config1_data
    from archi_b200 import B200VectorStore
    from archi_b200.embeddings import B200Embeddings
    from archi_b200.ingest import split_text
    from oracle import oracle as orc
    fx = config1_data.load()
    chunks, metas = [], []
    for doc in fx["docs"]:
        for i, c in enumerate(split_text(doc["text"], 1000, 0, "\n\n")):
            chunks.append(c)
            metas.append({"filename": doc["filename"], "chunk_index": i})
    queries = config1_data.queries(chunks, 20)
    ef = B200Embeddings(dtype="f32", seed=0, device=env.local_rank)
    name = "bench_config1"
    B200VectorStore.drop_collection(name, device=env.local_rank)
    store = B200VectorStore({}, ef, collection_name=name, device=env.local_rank)
    store.add_texts(chunks[:8], [dict(m) for m in metas[:8]])            # warm-up (cuDNN/cuBLAS handles, workspaces)
    B200VectorStore.drop_collection(name, device=env.local_rank)
    store = B200VectorStore({}, ef, collection_name=name, device=env.local_rank)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    store.add_texts(chunks, [dict(m) for m in metas])
    torch.cuda.synchronize()
    t_add = time.perf_counter() - t0
    store.similarity_search_with_score(queries[0], k=5)
    t0 = time.perf_counter()
    results = [store.similarity_search_with_score(q, k=5) for q in queries]
    t_search = time.perf_counter() - t0
    stored = store.native.read_rows(0, len(chunks))
    q_emb = np.asarray([ef.embed_query(q) for q in queries], dtype=np.float32)
    d_true, i_true = orc.c_scan_topk("cosine", stored, q_emb, 5)
    failed = 0
    for qi, res in enumerate(results):
        got = np.asarray([[chunks.index(d.page_content) for d, _ in res]])
        sc = np.asarray([[s for _, s in res]], dtype=np.float32)
        failed += 1 if orc.verify_topk("cosine", stored, q_emb[qi:qi + 1], 5, got, sc, 1e-5, d_true[qi:qi + 1], i_true[qi:qi + 1]) else 0
    # CPU leg: encoder forward in torch fp32 on the host cores (cpu_embed) + oracle.c scan
    model = copy.deepcopy(ef.model).to("cpu", torch.float32).eval()
    torch.set_num_threads(os.cpu_count() or 1)

    def cpu_embed(texts):
        outs = []
        texts = [t.replace("\n", " ") for t in texts]
        for s in range(0, len(texts), 32):
            ids, mask = ef.tokenizer(texts[s:s + 32], ef.max_seq_length)
            with torch.inference_mode():
                hidden = model(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask)).last_hidden_state
            outs.append(orc.c_pool_normalize(hidden.numpy(), mask))
        return np.concatenate(outs).astype(np.float32)
    t0 = time.perf_counter()
    cpu_rows = cpu_embed(chunks)
    t_cpu_embed = time.perf_counter() - t0
    t0 = time.perf_counter()
    for q in queries:
        orc.c_scan_topk("cosine", cpu_rows, cpu_embed([q]), 5)
    t_cpu_search = time.perf_counter() - t0
    B200VectorStore.drop_collection(name, device=env.local_rank)
    return {"config": {"workload": "configs[0]: archi docs/ chunked with the default data_manager config, MiniLM-L6 shape (random init, 384-d), top-5",
                       "rows": len(chunks), "dim": 384, "k": 5, "batch": 1, "storage": "f32", "metric": "cosine"},
            "files": fx["n_files"], "chunks": len(chunks),
            "ingest_chunks_per_s": len(chunks) / t_add, "similarity_search_queries_per_s": len(queries) / t_search,
            "parity_checked": len(queries), "parity_failed": failed,
            "max_abs_diff_vs_cpu_embeddings": float(np.abs(cpu_rows - stored).max()),
            "cpu": {"cores": os.cpu_count(), "kind": "port",
                    "cpu_embed_chunks_per_s": len(chunks) / t_cpu_embed,
                    "similarity_search_queries_per_s": len(queries) / t_cpu_search,
                    "what": "same weights, torch fp32 encoder forward on the host + oracle.c pool/normalise + oracle.c scan (embed_query + scan per query)"}}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--data", default="gaussian", choices=["gaussian", "latent", "dupes"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity", type=int, default=16, help="queries of the timed batch verified against oracle.c (0 = skip)")
    ap.add_argument("--sub-batches", default=None,
                    help="extra batch sizes reported under 'batches' (default: 1,64 at N=1, none at N>1)")
    ap.add_argument("--workloads", default=None,
                    help="sub-records: comma list of c3,c4,c5,c1,robust or 'none' (default: all when --workload is c2)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rows, dim, storage, k, dbatch, cfg_id, desc = WORKLOADS[args.workload]
    batch = args.batch or dbatch

    if args.impl == "reference":
        run_reference(args, rows, dim, storage, k, batch, desc, cfg_id)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)
    env = Env(world, rank, local_rank, dev, measured_peaks())

    if args.sub_batches is None:
        args.sub_batches = "1,64" if world == 1 else ""
    subs = [int(b) for b in args.sub_batches.split(",") if b]
    line = run_dense(env, args.workload, rows, dim, storage, k, batch, cfg_id, desc, args.steps, args.warmup,
                     data=args.data, sub_batches=subs, parity_n=args.parity)

    if args.workloads is None:
        args.workloads = "c3,c4,c5,c1,robust" if (args.workload == "c2" and args.data == "gaussian") else "none"
    wanted = [w for w in args.workloads.split(",") if w and w != "none"]
    workloads = {}
    sub_steps = max(5, min(args.steps, 20))
    for w in wanted:
        # every workload starts from an idle GPU, like the headline did: long tensor-core launches run into the
        # power cap within a step, and what the previous workload left in the power / thermal integrators would
        # otherwise decide the next one's clocks
        import torch as _torch
        _torch.cuda.synchronize()
        time.sleep(3.0)
        try:
            if w == "c3":
                r, d_, st_, k_, b_, cid, desc_ = WORKLOADS["c3"]
                rec = run_dense(env, "c3", r, d_, st_, k_, b_, cid, desc_, sub_steps, args.warmup, parity_n=args.parity)
            elif w == "c4":
                # configs[3] is 100M rows over 8 GPUs = 12.5M rows per GPU: the same per-GPU shard at every N
                # (at N = 8 this IS configs[3]); batch 1024 and batch 1
                r = 12_500_000 * world
                desc_ = (WORKLOADS["c4"][6] if world == 8 else
                         f"configs[3] at {world}/8 scale: {r // 1_000_000}M x 1024 bf16 chunks, 12.5M per GPU (the shard configs[3] puts on each of 8 GPUs), top-100")
                rec = run_dense(env, "c4s", r, 1024, "bf16", 100, 1024, 4, desc_, max(5, min(args.steps, 10)), args.warmup,
                                sub_batches=[1], parity_n=min(args.parity, 8), with_e2e=(world == 1))
                if rec is not None:
                    rec["scaling"] = "weak"
            elif w == "robust":
                rec = {}
                for data in ("latent", "dupes"):
                    r, d_, st_, k_, b_, cid, desc_ = WORKLOADS["c2"]
                    rr = run_dense(env, "c2", r, d_, st_, k_, b_, cid, desc_ + f" -- {data} rows instead of isotropic gaussian",
                                   sub_steps, args.warmup, data=data, parity_n=args.parity, with_e2e=False)
                    if rr is not None:
                        rec[data] = rr
                rec = rec or None
            elif w == "c5":
                rec = run_c5(env) if rank == 0 else None
                env.barrier()
            elif w == "c1":
                rec = run_c1(env) if rank == 0 else None
                env.barrier()
            else:
                continue
        except Exception as e:  # noqa: BLE001 - a failing sub-record must not lose the headline
            import traceback
            traceback.print_exc()
            rec = {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None
            try:
                env.barrier()
            except Exception:
                pass
        if rank == 0 and rec is not None:
            workloads[w] = rec

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            orc, corpus_h, q_h, threads, sample_q = cpu_reference_leg(rows, dim, k, target_s=12.0, cfg_id=cfg_id)
            t0 = time.perf_counter()
            orc.c_scan_topk("cosine", corpus_h, q_h[:sample_q], k, nthreads=threads, fast=True)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample_q / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                                    "sample": f"{sample_q} queries x {rows}x{dim} fp32 rows; {REF_NOTE}"}
            del corpus_h
            line["cpu_context"] = cpu_context(rows, dim, k, batch, cfg_id)
        else:
            line["cpu_baseline"] = None
        line["scaling"] = "strong"
        line["vs_baseline"] = None
        if workloads:
            line["workloads"] = workloads
        emit_json_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
