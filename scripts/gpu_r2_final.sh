#!/bin/bash
# Round 2, final evidence: the whole -m gpu suite (full-size parity included), the default bench line,
# the launch list with DRAM bytes (-> profiles/ncu_traffic.json), ncu --set full of the H+L and B2 launches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
what=${1:-all}
if [ "$what" = all ] || [ "$what" = tests ]; then
  timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -12 gpurun_out/r02_pytest_gpu_final.log
  head -2 gpurun_out/full_parity_spend256.log
fi
if [ "$what" = all ] || [ "$what" = bench ]; then
  time timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
  tail -4 gpurun_out/r02_bench_n1_final.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_final.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms/step %.1f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
print("roofline", {k: v for k, v in d["roofline"].items() if k in ("achieved", "frac", "traffic_over_algorithmic", "avg_launch_ms", "share_of_step", "fp_mul_per_s")})
for k, v in d.get("configs", {}).items():
    if k != "msm_sweep": print(k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a in ("proofs_per_s", "tx_per_s", "seconds", "oracle_sample")})
    else:
        for r in v: print("  msm 2^%d %s total %.1f acc %.1f ok=%s" % (r["log_n"], r["scalars"], r["ms_total"], r["ms_accumulate_kernel"], r["closed_form_ok"]))
print("circuit_path", {a: (round(b, 1) if isinstance(b, float) else b) for a, b in (d.get("circuit_path") or {}).items() if a in ("proofs_per_s_pipelined", "proofs_per_s_pipelined_with_self_check", "host_witness_per_s", "vs_synthetic_rows_e2e", "error")})
print("cpu", d.get("cpu_baseline"))
PY
  [ -s gpurun_out/r02_bench_reference_arm.json ] || timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
fi
B="python bench.py --steps 1 --warmup 1 --batch 64 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path"
if [ "$what" = all ] || [ "$what" = ncu ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -c 900 --csv \
     --log-file gpurun_out/r02_launches_final.csv $B > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/r02_launches_final.csv synth_ decode_ build_table pair_prep 2>/dev/null | head -14
  PYTHONPATH=. python scripts/traffic_from_launches.py gpurun_out/r02_launches_final.csv gpurun_out/ncu_traffic.json
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:msm_accumulate_g1 --launch-skip 2 -c 1 -o gpurun_out/r02_prof_final_acc_g1_hl -f $B > /dev/null 2>&1
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:msm_accumulate_g2 -c 1 -o gpurun_out/r02_prof_final_acc_g2 -f $B > /dev/null 2>&1
  ls -la gpurun_out/*.ncu-rep | tail -3
fi
