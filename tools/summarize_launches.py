"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (share of the step)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), us))
    tot = sum(u for _, u in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, u in rows:
        agg[n][0] += 1
        agg[n][1] += u
    print(f"# {path}: {len(rows)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'share':>7s}")
    for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:70s} {c:8d} {u:10.1f} {100*u/tot:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1])
