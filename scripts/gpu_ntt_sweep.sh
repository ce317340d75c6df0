#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/ntt_thin_sweep.jsonl
for v in 0 60 100 200; do
  MB200_NTT_THIN_KB=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-circuit-path --cpu-seconds 2 2>gpurun_out/thin_err.log | \
    python -c "import sys,json; l=json.loads(sys.stdin.readline()); print(json.dumps({'ntt_dyn_smem_kb': $v, 'value': round(l['value'],1), 'e2e': round(l['e2e']['value'],1), 'unpipelined_ms': round(l['device_ms_per_step_unpipelined'],1), 'parity': l['cpu_baseline']['gpu_proofs_byte_identical']}))" | tee -a gpurun_out/ntt_thin_sweep.jsonl
  tail -1 gpurun_out/thin_err.log
done
