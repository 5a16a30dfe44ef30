#!/usr/bin/env python
"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel: launches, total time, share of the sum.

    python tools/launch_summary.py gpurun_out/runNN/launches.csv > profiles/rNN_launches_summary.md

Per-launch times under ncu are serialised and cold-cache: the SHARES are what is comparable with bench.py's step.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0.0])
    unit = None
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        name = re.sub(r"^void |uwu::|\(anonymous namespace\)::|<unnamed>::", "", name)
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        agg[name][0] += 1
        agg[name][1] += v * scale
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# ncu launch list: {n} launches, {total / 1e3:.1f} ms summed kernel time (serialised, cold cache)\n")
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {c} | {t / 1e3:.2f} | {100 * t / total:.1f}% |")


if __name__ == "__main__":
    main()
