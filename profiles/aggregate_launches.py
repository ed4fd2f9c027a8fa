"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/aggregate_launches.py gpurun_out/launches.csv [steps]"""
import collections
import csv
import re
import sys


def main(path, steps=None):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-24s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-24s %6d %12.1f %10.1f %6.1f%%" % (k[:24], n, t, t / n, 100 * t / tot))
    print("total_us %.1f" % tot + ("  per_step_us %.1f" % (tot / steps) if steps else ""))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
