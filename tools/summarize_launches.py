#!/usr/bin/env python
"""Launch list -> markdown: aggregates an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` capture by kernel
name (count, total us, share, average us). Usage: python tools/summarize_launches.py gpurun_out/launches.csv [top_n]"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"<.*", "", name)          # drop template arguments
    name = re.sub(r"\(.*", "", name)         # and the parameter list
    return name.strip()[:90]


def main() -> None:
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]  # ncu banner lines
    for rec in csv.DictReader(lines):
        if rec.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(rec["Metric Value"].replace(",", ""))
        unit = (rec.get("Metric Unit") or "ns").lower()
        us = val / 1000.0 if unit in ("ns", "nsecond") else val * (1000.0 if unit.startswith("ms") else 1.0)
        rows.append((short(rec["Kernel Name"]), us))
    agg = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        agg[k][0] += 1
        agg[k][1] += us
    total = sum(v[1] for v in agg.values()) or 1.0
    print(f"{len(rows)} launches, {total / 1000.0:.2f} ms of kernel time\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{k}` | {n} | {us:.0f} | {100 * us / total:.1f}% | {us / n:.1f} |")


if __name__ == "__main__":
    main()
