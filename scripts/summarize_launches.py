"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean time, share."""
import collections
import csv
import sys


def main(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r["Kernel Name"], []).append(float(r["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    print("# %s: %d launches, %.3f ms total (ncu per-launch times are cold-cache and serialised: compare shares)" % (
        path, len(rows), tot / 1e6))
    print("%5s %12s %7s  %s" % ("count", "mean_us", "share%", "kernel"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%5d %12.1f %7.2f  %s" % (len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot, k[:150]))


if __name__ == "__main__":
    main(sys.argv[1])
