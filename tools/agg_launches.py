"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:72]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:3d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
