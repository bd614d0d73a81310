"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel (and per kernel x grid, which
separates the network's stages).  Per-launch times under ncu are cold-cache and serialised: compare SHARES.

    python tools/launch_summary.py gpurun_out/<tag>_launches_bench.csv [--by-grid] [--title "..."]
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("rb::<unnamed>::", "").replace("rb::", "")
    return name[:78]


def main():
    path = sys.argv[1]
    by_grid = "--by-grid" in sys.argv
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else path
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ki, gi, bi, vi, mi = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Block Size", "Metric Value", "Metric Name"))
    agg = defaultdict(lambda: [0.0, 0])
    total, n = 0.0, 0
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        ns = float(r[vi].replace(",", ""))
        key = short(r[ki]) + (("  grid " + r[gi] + " x " + r[bi]) if by_grid else "")
        agg[key][0] += ns
        agg[key][1] += 1
        total += ns
        n += 1
    print(title)
    print("per-launch times are cold-cache and serialised: compare SHARES. total %.3f ms over %d launches\n" % (total / 1e6, n))
    for key, (ns, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if ns / total < 0.002:
            continue
        print("%9.3f ms %5.1f%% %5d  %7.1f us/launch  %s" % (ns / 1e6, 100 * ns / total, cnt, ns / cnt / 1e3, key))


if __name__ == "__main__":
    main()
