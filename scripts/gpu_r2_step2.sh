#!/bin/bash
# Round 2, second GPU call: the multi-device C ABI, the shared-memory / TMA NTT as the default path, the new
# bench line (configs + MSM sweep), launch list with DRAM bytes per kernel, and where a 2^24 MSM spends its time.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
what=${1:-all}
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -14 gpurun_out/r02_pytest_gpu.log
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  ( time timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | tail -3
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f share %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["share_of_step"]))
    for k, v in d.get("configs", {}).items():
        print(k, json.dumps(v)[:600])
    print("circuit_path", json.dumps(d.get("circuit_path"))[:400])
    print("cpu", json.dumps(d.get("cpu_baseline"))[:300], d.get("parity_error"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r02_bench_n1.err").read()[-3000:])
PY
fi
if [ "$what" = ncu ] || [ "$what" = all ]; then
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
  B="python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-circuit-path --no-configs --no-msm-sweep"
  timeout 600 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv $B > /dev/null 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_msm24.csv \
     python scripts/msm_sweep.py --sizes 22 24 --kinds U --reps 1 > gpurun_out/r02_msm24_under_ncu.log 2>&1
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:msm_accumulate_g2 -c 1 -o gpurun_out/r02_prof_acc_g2 -f $B > /dev/null 2>&1
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:msm_accumulate_g1 --launch-skip 2 -c 1 -o gpurun_out/r02_prof_acc_g1_hl -f $B > /dev/null 2>&1
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:ntt_fused --launch-skip 1 -c 1 -o gpurun_out/r02_prof_ntt_fused -f $B > /dev/null 2>&1
  ls -la gpurun_out/ | tail -8
fi
if [ "$what" = msm ] || [ "$what" = all ]; then
  timeout 600 python scripts/msm_sweep.py --sizes 16 20 22 24 --reps 2 --check > gpurun_out/r02_msm_sweep_before.jsonl 2>&1; tail -8 gpurun_out/r02_msm_sweep_before.jsonl
fi
