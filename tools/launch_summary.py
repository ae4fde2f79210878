#!/usr/bin/env python
"""Summarises an ncu launch list (the `--metrics gpu__time_duration.sum --clock-control none --csv` pass of
/opt/skills/guides/B200_PROFILING.md over ONE bench.py step) into per-kernel totals and shares.
Usage: python tools/launch_summary.py profiles/rNN_launches.csv [out.txt]

ncu serialises the launches and measures each one cold, so the absolute times are not the in-step times: the point of
this file is each kernel's SHARE of the step, to set beside bench.py's live CUDA-event numbers."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        rows.append((name, ns))
    total = sum(ns for _, ns in rows)
    agg = collections.OrderedDict()
    for name, ns in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    out = [f"total {total / 1e6:.1f} ms over {len(rows)} launches (ncu per-launch times: cold-cache, serialised; compare SHARES)"]
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{ns / 1e6:9.2f} ms {100 * ns / total:5.1f}%  n={n:4d}  avg {ns / n / 1e3:9.1f} us  {name}")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
