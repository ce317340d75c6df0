#!/bin/bash
# Round 2, fourth GPU call: SIMD witness generator on the box's CPU, folded digit walk (c_HL 16 vs 17),
# block scans / warp-aggregated ordering (launch list), inlined Horner + inversion (MSM sweep).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
what=${1:-all}
lscpu | grep -E "Model name|^CPU\(s\)" ; lscpu | grep -o -E "avx512ifma" | sort -u
if [ "$what" = all ] || [ "$what" = tests ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02_pytest_gpu_step4.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu_step4.log
fi
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f share %.3f knobs %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["config"]["knobs"]))
        cp = d.get("circuit_path") or {}
        print("   circuit_path", {a: (round(b, 1) if isinstance(b, float) else b) for a, b in cp.items() if a in ("proofs_per_s_pipelined", "host_witness_per_s", "host_threads", "vs_synthetic_rows_e2e", "error")})'
if [ "$what" = all ] || [ "$what" = ab ]; then
  for c in 16 17; do
    MB200_C_HL=$c timeout 400 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline > gpurun_out/r02_ab_chl$c.json 2> gpurun_out/r02_ab_chl$c.err
    python -c "$P" < gpurun_out/r02_ab_chl$c.json; tail -2 gpurun_out/r02_ab_chl$c.err
  done
fi
if [ "$what" = all ] || [ "$what" = launches ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -c 900 --csv \
     --log-file gpurun_out/r02_launches_step4.csv python bench.py --steps 1 --warmup 1 --batch 64 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/r02_launches_step4.csv synth_ decode_ build_table | head -28
fi
if [ "$what" = all ] || [ "$what" = msm ]; then
  timeout 600 python scripts/msm_sweep.py --sizes 16 18 20 22 24 --reps 3 > gpurun_out/r02_msm_sweep_n1_step4.jsonl 2> gpurun_out/r02_msm_sweep_n1_step4.err
  python - <<'PY'
import json
for l in open("gpurun_out/r02_msm_sweep_n1_step4.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["log_n"], d["scalars"], "total %.1f acc %.1f limit %.1f %s ok=%s" % (d["ms_total"], d["ms_accumulate_kernel"], d["limit_ms_1.3x_acc_plus_5"], d["within_limit"], d["closed_form_ok"]))
PY
  tail -3 gpurun_out/r02_msm_sweep_n1_step4.err
fi
