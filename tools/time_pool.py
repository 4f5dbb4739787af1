"""Device time of the pool+normalise kernels (ring kernel vs one CTA per sequence) on encoder-shaped inputs.

    python tools/time_pool.py [--json out.json]

Each variant is timed as a CUDA graph of 20 launches (no host time between launches) and as a plain loop of
Python calls; bytes = live tokens x H x elt + mask + outputs.  ARCHI_POOL_RING selects the kernel."""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from archi_b200 import _native as N
from archi_b200.store import _current_stream_ptr, _torch_dtype_code


def launch(h, m, o32, o16):
    B, L, H = h.shape
    N.check(N.lib().archi_pool_normalize(
        ctypes.c_void_p(h.data_ptr()), _torch_dtype_code(h), ctypes.c_void_p(m.data_ptr()), _torch_dtype_code(m),
        B, L, H, ctypes.c_void_p(o16.data_ptr()), ctypes.c_void_p(o32.data_ptr()),
        ctypes.c_void_p(_current_stream_ptr(h.device.index))))


def time_variant(h, m, reps=20):
    B, L, H = h.shape
    o32 = torch.empty((B, H), dtype=torch.float32, device=h.device)
    o16 = torch.empty((B, H), dtype=torch.bfloat16, device=h.device)
    for _ in range(3):
        launch(h, m, o32, o16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch(h, m, o32, o16)
    e1.record()
    torch.cuda.synchronize()
    loop_ms = e0.elapsed_time(e1) / reps
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        launch(h, m, o32, o16)
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                launch(h, m, o32, o16)
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return loop_ms, best, o32


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--ring", default="both", choices=["0", "1", "both"])
    ap.add_argument("--cases", type=int, default=99)
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6545.9))
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    out = []
    for (B, L, H, dt) in [(1024, 256, 384, torch.bfloat16), (4096, 256, 384, torch.bfloat16), (1024, 512, 768, torch.bfloat16),
                          (1024, 256, 1024, torch.float32), (512, 256, 384, torch.bfloat16),
                          (300, 128, 384, torch.bfloat16)][:args.cases]:
        h = torch.randn((B, L, H), device=dev, dtype=torch.float32).to(dt)
        lens = torch.randint(L // 2, L + 1, (B,), device=dev)
        m = (torch.arange(L, device=dev)[None, :] < lens[:, None]).to(torch.int64)
        live = int(m.sum().item())
        nbytes = live * H * h.element_size() + B * L * 8 + B * H * 6
        rec = {"B": B, "L": L, "H": H, "dtype": str(dt).replace("torch.", ""), "algorithmic_bytes": nbytes}
        ref = None
        for ring in (("0", "1") if args.ring == "both" else (args.ring,)):
            os.environ["ARCHI_POOL_RING"] = ring
            loop_ms, graph_ms, o = time_variant(h, m)
            name = "ring" if ring == "1" else "cta_per_sequence"
            rec[name] = {"loop_ms": loop_ms, "graph_ms": graph_ms, "GBps": nbytes / graph_ms * 1e-6,
                         "frac_hbm": nbytes / graph_ms * 1e-6 / hbm}
            if ref is None:
                ref = o.clone()
            else:
                rec["max_abs_diff_between_kernels"] = float((o - ref).abs().max().item())
        del os.environ["ARCHI_POOL_RING"]
        out.append(rec)
        print(json.dumps(rec), flush=True)
    if args.json:
        json.dump({"hbm_peak_GBps": hbm, "cases": out}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
