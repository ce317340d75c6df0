#!/bin/bash
# Round 2, fifth GPU call: H+L base ranges (DRAM re-reads of the table), direct bucket writes, MSM tails.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
what=${1:-all}
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f share %.3f knobs %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["config"]["knobs"]))'
if [ "$what" = all ] || [ "$what" = tests ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "msm or prove or rows_kernel" > gpurun_out/r02_pytest_gpu_step5.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_step5.log
  MB200_HL_SLABS=4 timeout 600 python -m pytest tests -m gpu -x -q -k "prove" 2>&1 | tail -2
fi
if [ "$what" = all ] || [ "$what" = ab ]; then
  for sl in 1 4 8 16; do
    MB200_HL_SLABS=$sl timeout 400 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/r02_ab_hlslabs$sl.json 2> gpurun_out/r02_ab_hlslabs$sl.err
    python -c "$P" < gpurun_out/r02_ab_hlslabs$sl.json; tail -2 gpurun_out/r02_ab_hlslabs$sl.err
  done
fi
if [ "$what" = all ] || [ "$what" = launches ]; then
  MB200_HL_SLABS=8 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -c 1400 --csv \
     --log-file gpurun_out/r02_launches_step5_slabs8.csv python bench.py --steps 1 --warmup 1 --batch 64 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/r02_launches_step5_slabs8.csv synth_ decode_ build_table pair_prep 2>/dev/null | head -16
fi
if [ "$what" = all ] || [ "$what" = msm ]; then
  timeout 600 python scripts/msm_sweep.py --sizes 16 18 20 22 24 --reps 3 > gpurun_out/r02_msm_sweep_n1_step5.jsonl 2> gpurun_out/r02_msm_sweep_n1_step5.err
  python - <<'PY'
import json
for l in open("gpurun_out/r02_msm_sweep_n1_step5.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["log_n"], d["scalars"], "total %.1f acc %.1f limit %.1f %s ok=%s" % (d["ms_total"], d["ms_accumulate_kernel"], d["limit_ms_1.3x_acc_plus_5"], d["within_limit"], d["closed_form_ok"]))
PY
  tail -3 gpurun_out/r02_msm_sweep_n1_step5.err
fi
