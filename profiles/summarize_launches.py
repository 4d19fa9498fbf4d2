"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share, average."""
import collections
import csv
import sys


def main(path, title=""):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    if title:
        print("# " + title)
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
    print(f"{'kernel':72s} {'n':>5s} {'total_us':>11s} {'share':>7s} {'avg_us':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} {v[0]:5d} {v[1]:11.1f} {v[1] / tot:7.3f} {v[1] / v[0]:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
