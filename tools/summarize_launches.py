#!/usr/bin/env python
"""ncu --csv launch list (gpu__time_duration.sum) -> per-kernel table (count, total, share)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        rows.append((name, ns))
    tot = sum(ns for _, ns in rows) or 1.0
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    lines = ["| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{n}` | {c} | {ns / 1e6:.2f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.1f} |")
    lines.append(f"| **total** | {len(rows)} | {tot / 1e6:.2f} | 100% | |")
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
