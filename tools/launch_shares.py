"""Per-kernel device time of an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file ...).

    python tools/launch_shares.py gpurun_out/r02_launches_c2.csv [searches]

Prints, for the LAST `searches` searches in the list (default 2: the timed steps of the profiled command), every
launch with its duration, and the share of each kernel in one search.  ncu serialises the launches and runs them
cold-cache, so only the SHARES are comparable with the bench line, not the absolute times."""
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*$", "", name.replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", ""))
    return name.replace("void ", "").replace("archi::", "").replace("__nv_bfloat16", "bf16")


def main():
    path = sys.argv[1]
    n_search = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
            rows.append((short(r["Kernel Name"]), v, r.get("Grid Size", ""), r.get("Block Size", "")))
    # a search = the launches from one tc_prep_kernel (or scan_topk for single queries) to the next
    starts = [i for i, r in enumerate(rows) if r[0].startswith("tc::tc_prep_kernel")]
    if not starts:
        starts = [i for i, r in enumerate(rows) if r[0].startswith("scan_topk_kernel")]
    print(f"{len(rows)} launches of archi kernels, {len(starts)} searches in the list")
    for s_i in range(max(0, len(starts) - n_search), len(starts)):
        a = starts[s_i]
        b = starts[s_i + 1] if s_i + 1 < len(starts) else len(rows)
        seg = rows[a:b]
        tot = sum(r[1] for r in seg)
        print(f"\nsearch {s_i}: {len(seg)} launches, {tot:.1f} us of kernel time")
        for name, us, grid, block in seg:
            print(f"  {us:10.1f} us  {us / tot * 100:5.1f} %  {name[:70]:70s} grid {grid} block {block}")


if __name__ == "__main__":
    main()
