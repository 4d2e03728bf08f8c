#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and mean
duration, share of the summed GPU time.  Usage: tools/launch_shares.py launches.csv [out.txt] ["note"]"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"void |unnamed>::|dvt::|\(anonymous namespace\)::|<unnamed>::", "", name)
    m = re.match(r"([A-Za-z0-9_]+)(<[^>]*>)?", name.strip())
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = short(r[4])
        agg[k][0] += 1
        agg[k][1] += float(r[-1])
    total = sum(v[1] for v in agg.values())
    lines = [f"# per-kernel totals of {path} ({len(rows)} launches; durations are ncu-serialised, cold-cache: compare SHARES)"]
    if len(sys.argv) > 3:
        lines.append("# " + sys.argv[3])
    lines.append(f"{'kernel':<58}{'launches':>9}{'total us':>12}{'mean us':>10}{'share':>8}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:<58}{n:>9}{t / 1e3:>12.1f}{t / n / 1e3:>10.2f}{100 * t / total:>7.1f}%")
    lines.append(f"{'TOTAL':<58}{len(rows):>9}{total / 1e3:>12.1f}")
    txt = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
