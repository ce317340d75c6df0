#!/usr/bin/env python
"""profiles/ncu_traffic.json from an ncu metrics pass of one 64-proof Spend chunk.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
      --clock-control none --csv --log-file L.csv python bench.py --steps 1 --warmup 1 --batch 64 ...
  python scripts/traffic_from_launches.py L.csv profiles/ncu_traffic.json

The four accumulate launches of the LAST chunk in the list are the four queries: the G2 launch is B2; of
the three G1 launches the one with the most DRAM traffic is H+L, then A, then B1.  Algorithmic bytes are
n_proofs x n_bases x 128 (G1) or x 224 (G2), SURVEY.md §8(d), for the Spend shape.
"""
import csv
import json
import sys

from masp_b200 import synthetic as syn

UNIT = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    n_proofs = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    rows = {}
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = rows.setdefault(k, {"name": r["Kernel Name"].split("(")[0].replace("mb::", "")})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    seq = [rows[k] for k in sorted(rows)]
    acc = [d for d in seq if d["name"].startswith("msm_accumulate_g")]
    last = acc[-4:]
    g2 = [d for d in last if d["name"] == "msm_accumulate_g2"]
    g1 = sorted([d for d in last if d["name"] == "msm_accumulate_g1"],
                key=lambda d: -(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]))
    assert len(g2) == 1 and len(g1) == 3, [d["name"] for d in last]
    sh = syn.SPEND
    bases = {"H+L": sh.h_len + sh.n_aux, "A": sh.a_len, "B1": sh.b_len, "B2": sh.b_len}
    out = {}
    for q, d, per in (("A", g1[1], 128), ("B1", g1[2], 128), ("H+L", g1[0], 128), ("B2", g2[0], 224)):
        traffic = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        alg = n_proofs * bases[q] * per
        out[q] = {"kernel": d["name"], "dram_bytes": int(traffic), "algorithmic_bytes": alg,
                  "ms": round(d["gpu__time_duration.sum"], 3), "l2_hit_pct": round(d.get("lts__t_sector_hit_rate.pct", 0.0), 1),
                  "traffic_over_algorithmic": round(traffic / alg, 3)}
    tot = sum(v["dram_bytes"] for v in out.values())
    alg = sum(v["algorithmic_bytes"] for v in out.values())
    doc = {
        "source": "ncu metrics pass on a B200 (--clock-control none), one %d-proof Spend chunk, the four accumulate launches "
                  "of the last chunk in %s (scripts/traffic_from_launches.py)" % (n_proofs, src),
        "by_query": out,
        "dram_bytes_per_chunk_all_four_launches": tot, "algorithmic_bytes_per_chunk_all_four_launches": alg,
        "dram_bytes_per_launch_avg": tot / 4.0, "algorithmic_bytes_per_launch_avg": alg / 4.0,
        "traffic_over_algorithmic": round(tot / alg, 3),
        "note": "H+L: every wave of resident blocks (444 of ~31 000) walks its buckets' whole entry lists, i.e. the whole 356 MB "
                "window table: ~70 waves x 356 MB.  ~6 % of the HBM peak while the multiplier is 84 % busy -- not what bounds it; "
                "cutting the query into L2-sized base ranges removes most of it and is slower (DESIGN.md §7b)",
    }
    with open(dst, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps({q: (v["ms"], v["traffic_over_algorithmic"]) for q, v in out.items()}), doc["traffic_over_algorithmic"])


if __name__ == "__main__":
    main()
