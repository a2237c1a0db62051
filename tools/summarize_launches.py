#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total time and share.
usage: python tools/summarize_launches.py profiles/<file>.csv"""
import collections
import csv
import sys


def summarize(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = ["| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / tot * 100:.1f}% | {v[1] / v[0] / 1e3:.1f} |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
