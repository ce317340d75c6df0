#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms %.1f share %.3f acc_ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["roofline"]["avg_launch_ms"]))
    else: print(l, end="")'
timeout 300 python -m pytest tests -m gpu -x -q -k "prove_spend or msm_g1 or msm_g2 or heavy" 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline $BENCH_ARGS 2>&1 | python -c "$P"
