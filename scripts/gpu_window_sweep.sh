#!/bin/bash
# window-size sweep: proofs/s of the headline bench for a few (c_HL, c_A/B1/B2) choices
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/window_sweep.jsonl
for cfg in "16 14 14 14" "15 14 14 14" "14 14 14 14" "16 13 13 13" "16 12 12 12" "16 13 13 12" "15 13 13 13" "15 12 12 12"; do
  set -- $cfg
  MB200_C_HL=$1 MB200_C_A=$2 MB200_C_B1=$3 MB200_C_B2=$4 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-circuit-path 2>/dev/null | \
    python -c "import sys,json; l=json.loads(sys.stdin.readline()); print(json.dumps({'c': '$cfg', 'value': round(l['value'],1), 'e2e': round(l['e2e']['value'],1), 'unpipelined_ms': round(l['device_ms_per_step_unpipelined'],1), 'acc_share': round(l['roofline']['share_of_step'],3)}))" | tee -a gpurun_out/window_sweep.jsonl
done
