"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (shares of one step; ncu times are
cold-cache and serialised).  Usage: python tools/launch_summary.py launches.csv [--list PATTERN]"""
import csv
import sys


def main(path, pattern=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt, seq = {}, {}, []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0][:70]
        v, u = float(r[mv].replace(",", "")), r[mu]
        ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
        tot[name] = tot.get(name, 0.0) + ms
        cnt[name] = cnt.get(name, 0) + 1
        seq.append((name, ms))
    total = sum(tot.values())
    print(f"# {path}: {len(seq)} launches, {total:.3f} ms of kernel time under ncu")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v:9.3f} ms {100 * v / total:5.1f} % x{cnt[k]:4d}  {k}")
    if pattern:
        print(f"# launches matching {pattern!r}, in order")
        for i, (n, ms) in enumerate(x for x in seq if pattern in x[0]):
            print(f"{i:4d} {ms:8.4f} ms  {n}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--list" else None)
