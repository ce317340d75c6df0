#!/bin/bash
# Round 2, sixth GPU call: register-capped accumulate kernels (13 / 14 / 15 warps per SM) against the default.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f share %.3f acc_ms %.2f knobs %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["roofline"]["avg_launch_ms"], d["config"]["knobs"]))'
MB200_ACC_REGS=136 timeout 300 python -m pytest tests -m gpu -x -q -k "msm_g1 or prove_spend or heavy" 2>&1 | tail -2
for r in 0 168 152 144 136; do
  MB200_ACC_REGS=$r timeout 400 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/r02_ab_accregs$r.json 2> gpurun_out/r02_ab_accregs$r.err
  python -c "$P" < gpurun_out/r02_ab_accregs$r.json; tail -2 gpurun_out/r02_ab_accregs$r.err
done
