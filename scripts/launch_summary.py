#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum,
lts__t_sector_hit_rate.pct] --csv` launch list by kernel: launches, total / average duration and share,
and -- when the capture has them -- DRAM bytes per launch and the L2 hit rate.

  python scripts/launch_summary.py gpurun_out/r02_launches.csv [kernel names to leave out ...]
"""
import collections
import csv
import sys

UNIT = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ii, ki, mi, vi, ui = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launches = collections.OrderedDict()
    for row in r:
        if len(row) <= vi:
            continue
        try:
            v = float(row[vi].replace(",", ""))
        except ValueError:
            continue
        v *= UNIT.get(row[ui], 1.0)
        launches.setdefault(row[ii], {"kernel": row[ki].split("(")[0]})[row[mi]] = v
    return list(launches.values())


def main():
    rows = load(sys.argv[1])
    skip = set(sys.argv[2:])  # kernel names to leave out (one-time setup kernels)
    agg = collections.OrderedDict()
    for l in rows:
        if l["kernel"] in skip or "gpu__time_duration.sum" not in l:
            continue
        a = agg.setdefault(l["kernel"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "hit": 0.0, "hit_n": 0})
        a["n"] += 1
        a["us"] += l["gpu__time_duration.sum"]
        a["rd"] += l.get("dram__bytes_read.sum", 0.0)
        a["wr"] += l.get("dram__bytes_write.sum", 0.0)
        if "lts__t_sector_hit_rate.pct" in l:
            a["hit"] += l["lts__t_sector_hit_rate.pct"]
            a["hit_n"] += 1
    total = sum(a["us"] for a in agg.values())
    print("%d launches, %.1f ms total (cold-cache, serialised: compare shares)" % (sum(a["n"] for a in agg.values()), total / 1e3))
    have_mem = any(a["rd"] or a["wr"] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        line = "%-26s n=%4d total %10.1f us %6.1f%%  avg %9.1f us" % (k, a["n"], a["us"], 100 * a["us"] / total, a["us"] / a["n"])
        if have_mem:
            line += "  dram rd %8.1f MB wr %8.1f MB /launch" % (a["rd"] / a["n"] / 1e6, a["wr"] / a["n"] / 1e6)
            if a["hit_n"]:
                line += "  L2 hit %5.1f%%" % (a["hit"] / a["hit_n"])
        print(line)


if __name__ == "__main__":
    main()
