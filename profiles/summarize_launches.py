#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python profiles/summarize_launches.py gpurun_out/launches.csv STEPS
STEPS = number of bench steps the capture holds (warm-up + timed + e2e legs)."""
import collections
import csv
import re
import sys


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1e-3)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"<.*", "", name)[:70]
        if name.strip() in ("void at::", "void", "void at::native::"):
            name = row["Kernel Name"][:110]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"| kernel | launches/step | us/step | share |\n|---|---|---|---|")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v / tot < 0.002:
            continue
        print(f"| `{n}` | {c / steps:.1f} | {v / steps:.1f} | {100 * v / tot:.1f}% |")
    print(f"\nsum of kernel durations: {tot / steps / 1000:.3f} ms/step over {steps} steps")


if __name__ == "__main__":
    main()
