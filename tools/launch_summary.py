"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, time and share.
usage: python tools/launch_summary.py <launches.csv> [steps] [top]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        us = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
        k = row["Kernel Name"][:100]
        agg[k][0] += 1
        agg[k][1] += us
        tot += us
    own = sum(t for k, (c, t) in agg.items() if "tagrec::" in k)
    print(f"{sum(c for c, _ in agg.values()) / steps:.0f} launches / step, {tot / steps / 1000:.2f} ms kernel time / step, "
          f"tagrec:: share {100 * own / tot:.0f} %")
    print("| kernel | launches / step | µs / step | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"| `{k[:72]}` | {c / steps:g} | {t / steps:.1f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
