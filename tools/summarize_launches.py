"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/r01_launches_summary.md
"""
import csv
import io
import re
import sys
from collections import defaultdict


def main(path):
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = defaultdict(lambda: [0, 0.0, 1e30, 0.0])
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        a = agg[name]
        a[0] += 1
        a[1] += ns
        a[2] = min(a[2], ns)
        a[3] = max(a[3], ns)
    tot = sum(a[1] for a in agg.values())
    print(f"launches: {sum(a[0] for a in agg.values())}, total device time {tot / 1e3:.1f} us "
          "(ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)\n")
    print("| kernel | launches | total us | share | min us | max us |")
    print("|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e3:.1f} | {a[3] / 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
