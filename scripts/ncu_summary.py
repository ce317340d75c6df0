#!/usr/bin/env python
"""Key metrics of an `ncu --set full` capture -> markdown + profiles/ncu_traffic.json.

  python scripts/ncu_summary.py gpurun_out/prof_acc_g1.ncu-rep profiles/r01_ncu_msm_accumulate_g1.md
"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, % of 64 warps"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per warp instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy, %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe (IMAD.WIDE) busy, %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipes (heavy+lite) busy, %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe busy, %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe, %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput, % of peak"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate, %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate, %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps per scheduler per cycle"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait (fixed-latency dependency)"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall samples: math pipe throttle"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard (memory)"),
    ("smsp__pcsamp_warps_issue_stalled_no_instructions", "stall samples: no instructions (i-cache)"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall samples: not selected"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "stall samples: selected (issuing)"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full --clock-control none: %s" % os.path.basename(rep), ""]
    traffic = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        lines += ["## launch %s: `%s`" % (r[col["ID"]], name), "", "| metric | value | unit |", "|---|---|---|"]
        for key, label in WANT:
            if key in col:
                lines.append("| %s (`%s`) | %s | %s |" % (label, key, r[col[key]], units[col[key]]))
        lines.append("")
        try:
            def to_bytes(k):
                v, u = float(r[col[k]].replace(",", "")), units[col[k]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic[name + "_dram_bytes_per_launch"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            traffic[name + "_grid"] = r[col["launch__grid_size"]]
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, {k: v for k, v in traffic.items()})


if __name__ == "__main__":
    main()
