"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py <launches.csv> <steps_in_run> [--grid]"""
import collections
import csv
import re
import sys


def main():
    path, steps = sys.argv[1], float(sys.argv[2])
    by_grid = "--grid" in sys.argv
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("marlc::", "")
        name = re.sub(r"<.*", lambda m: m.group(0)[:28], name)
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        key = (name, row["Grid Size"]) if by_grid else name
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {tot / steps:.1f} us of kernel time per step ({steps:g} steps in the run)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{str(k)[:70]:70s} n/step={v[0] / steps:7.1f} us/step={v[1] / steps:9.1f} avg={v[1] / v[0]:8.2f}us share={v[1] / tot * 100:5.1f}%")


if __name__ == "__main__":
    main()
