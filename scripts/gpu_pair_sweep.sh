#!/bin/bash
# batched-affine pair rounds in front of the accumulation: parity, then proofs/s per setting
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
: > gpurun_out/pair_sweep.jsonl
for cfg in ${PAIR_CFGS:-"0 16 32" "1 16 32" "2 16 32" "3 16 32" "4 16 32" "3 16 64" "3 16 128" "3 32 64"}; do
  set -- $cfg
  MB200_PAIR_ROUNDS=$1 MB200_PASS_INSTANCES=$2 MB200_PAIR_B=$3 timeout 300 python bench.py --steps 3 --warmup 3 --no-circuit-path --cpu-seconds 2 2>gpurun_out/pair_err.log | \
    python -c "import sys,json; l=json.loads(sys.stdin.readline()); print(json.dumps({'rounds_pass_B': '$cfg', 'value': round(l['value'],1), 'e2e': round(l['e2e']['value'],1), 'unpipelined_ms': round(l['device_ms_per_step_unpipelined'],1), 'acc_share': round(l['roofline']['share_of_step'],3), 'parity': l['cpu_baseline']['gpu_proofs_byte_identical'], 'power_w': l['clocks'].get('power_w_max')}))" | tee -a gpurun_out/pair_sweep.jsonl
  tail -2 gpurun_out/pair_err.log
done
