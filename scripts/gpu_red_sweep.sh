#!/bin/bash
# reduction fan-in sweep: proofs/s of the headline bench for (log T level 0, log T upper levels)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/red_sweep.jsonl
for cfg in "4 4" "4 3" "4 2" "5 3" "5 2" "3 3" "3 2" "6 2"; do
  set -- $cfg
  MB200_RED_LOG_T0=$1 MB200_RED_LOG_T1=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-circuit-path 2>/dev/null | \
    python -c "import sys,json; l=json.loads(sys.stdin.readline()); print(json.dumps({'log_t0_t1': '$cfg', 'value': round(l['value'],1), 'e2e': round(l['e2e']['value'],1), 'unpipelined_ms': round(l['device_ms_per_step_unpipelined'],1), 'acc_share': round(l['roofline']['share_of_step'],3)}))" | tee -a gpurun_out/red_sweep.jsonl
done
