#!/bin/bash
# Round 2, eleventh GPU call: Fr reduction rows as IMAD.WIDE (opaque INV), first-row products as mul.wide.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f share %.3f acc_ms %.2f fpmul %.4g" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["fp_mul_per_s"]))'
timeout 500 python -m pytest tests -m gpu -x -q -k "selftest or ntt or h_coeff or msm or prove or verify or rows_kernel" 2>&1 | tail -2
for i in 1 2; do
timeout 400 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/r02_ab_wide_$i.json 2> gpurun_out/r02_ab_wide_$i.err
python -c "$P" < gpurun_out/r02_ab_wide_$i.json; tail -2 gpurun_out/r02_ab_wide_$i.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_step11.csv python bench.py --steps 1 --warmup 1 --batch 64 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_step11.csv synth_ decode_ build_table pair_prep 2>/dev/null | head -9
