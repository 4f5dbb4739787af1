"""Markdown digest of an .ncu-rep (ncu --set full [--import-source on]): one metric table per captured launch and,
when the report carries source counters, the SASS lines with the most stall samples.

    python tools/ncu_digest.py gpurun_out/x.ncu-rep [--top 12] [--max-launches 4] > profiles/x.md

Runs wherever ncu is installed (no GPU needed)."""
import argparse
import csv
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active, % of active cycles"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_elapsed.avg", "SM elapsed cycles (avg)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, % of peak warps"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--max-launches", type=int, default=4)
    a = ap.parse_args()
    rows = list(csv.reader(run([a.report, "--page", "raw", "--csv"]).splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    print(f"# ncu digest of `{a.report.split('/')[-1]}` ({len(body)} launch(es) captured)\n")
    for n, r in enumerate(body[:a.max_launches]):
        print(f"## launch {n}: `{r[hdr.index('Kernel Name')][:110]}`\n")
        print("| metric | value |\n|---|---|")
        for key, label in METRICS:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        print()
    src = list(csv.reader(run([a.report, "--page", "source", "--csv"]).splitlines()))
    secs, cur = [], None
    for r in src:
        if "Source" in r and "# Samples" in r:
            cur = {"hdr": r, "body": []}
            secs.append(cur)
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["body"].append(r)
    seen = 0
    for s in secs:
        h, b = s["hdr"], s["body"]
        si, sa, ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        tot = sum(float(r[sa] or 0) for r in b)
        if tot <= 0 or seen >= a.max_launches:
            continue
        print(f"## SASS lines with the most stall samples, launch {seen} ({tot:.0f} samples)\n")
        print("| share | executed | instruction | top stall reasons |\n|---|---|---|---|")
        for r in sorted(b, key=lambda x: -float(x[sa] or 0))[:a.top]:
            why = sorted(((float(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:2]
            print(f"| {float(r[sa]) / tot * 100:.1f} % | {r[ex]} | `{r[si].strip()[:90]}` | "
                  + ", ".join(f"{w} {v:.0f}" for v, w in why if v > 0) + " |")
        print()
        seen += 1


if __name__ == "__main__":
    main()
