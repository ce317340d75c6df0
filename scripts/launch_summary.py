#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    out = []
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "nsecond": 1e-3}.get(row[ui], 1.0)
        out.append((row[ki].split("(")[0], v))
    return out


def main():
    rows = load(sys.argv[1])
    skip = set(sys.argv[2:])  # kernel names to leave out (one-time setup kernels)
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for n, v in rows:
        if n in skip:
            continue
        tot[n] += v
        cnt[n] += 1
    T = sum(tot.values())
    print("%d launches, %.1f ms total (cold-cache, serialised: compare shares)" % (sum(cnt.values()), T / 1e3))
    for n, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("%-24s n=%4d total %10.1f us  %5.1f%%  avg %9.1f us" % (n[:24], cnt[n], v, 100 * v / T, v / cnt[n]))


if __name__ == "__main__":
    main()
