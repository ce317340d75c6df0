#!/bin/bash
# Round 2, seventh GPU call: Y3 of the G1 mixed addition through one fused reduction (sop2).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f share %.3f acc_ms %.2f fpmul %.3g knobs %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["fp_mul_per_s"], d["config"]["knobs"]))'
timeout 400 python -m pytest tests -m gpu -x -q -k "selftest or msm or prove or special" 2>&1 | tail -2
for i in 1 2; do
timeout 400 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/r02_ab_sop2_$i.json 2> gpurun_out/r02_ab_sop2_$i.err
python -c "$P" < gpurun_out/r02_ab_sop2_$i.json; tail -2 gpurun_out/r02_ab_sop2_$i.err
done
