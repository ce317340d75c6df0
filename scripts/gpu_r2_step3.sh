#!/bin/bash
# Round 2, third GPU call: tree combine + slab pipeline of the standalone MSM, full-size parity tests.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
what=${1:-all}
if [ "$what" = msm ] || [ "$what" = all ]; then
  timeout 300 python -m pytest tests -m gpu -x -q -k "msm or prove_tiny or prove_spend" 2>&1 | tail -3
  timeout 900 python scripts/msm_sweep.py --sizes 16 18 20 22 24 --reps 3 > gpurun_out/r02_msm_sweep_n1.jsonl 2> gpurun_out/r02_msm_sweep_n1.err
  python - <<'PY'
import json
for l in open("gpurun_out/r02_msm_sweep_n1.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["log_n"], d["scalars"], "total %.1f acc %.1f limit %.1f %s ok=%s" % (d["ms_total"], d["ms_accumulate_kernel"], d["limit_ms_1.3x_acc_plus_5"], d["within_limit"], d["closed_form_ok"]))
PY
  tail -5 gpurun_out/r02_msm_sweep_n1.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_msm_after.csv \
     python scripts/msm_sweep.py --sizes 16 24 --kinds U --reps 1 --no-check > /dev/null 2>&1
  python scripts/launch_summary.py gpurun_out/r02_launches_msm_after.csv synth_g1 decode_g1 | head -24
fi
if [ "$what" = parity ] || [ "$what" = all ]; then
  timeout 1500 python -m pytest tests/test_zz_full_size_parity.py -m gpu -q --durations=6 > gpurun_out/r02_pytest_full_parity.log 2>&1; tail -12 gpurun_out/r02_pytest_full_parity.log
  head -3 gpurun_out/full_parity_spend256.log
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  timeout 600 python bench.py --steps 4 --warmup 3 --no-msm-sweep > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n1_b.json").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f share %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["share_of_step"]))
    for k, v in d.get("configs", {}).items():
        print(k, json.dumps(v)[:300])
    print("circuit_path", json.dumps(d.get("circuit_path"))[:500])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r02_bench_n1_b.err").read()[-3000:])
PY
fi
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core" ; lscpu | grep -o -E "avx512ifma|adx|bmi2" | sort -u | tr '\n' ' '; free -g | head -2
