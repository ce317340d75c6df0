#!/bin/bash
# in-flight chunk contexts / chunk size sweep of the headline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/ctx_sweep.jsonl
for cfg in "64 3" "64 4" "64 5" "32 6" "32 4" "128 3" "128 2"; do
  set -- $cfg
  timeout 300 python bench.py --steps 3 --warmup 3 --chunk $1 --streams $2 --no-cpu-baseline --no-circuit-path 2>/dev/null | \
    python -c "import sys,json; l=json.loads(sys.stdin.readline()); print(json.dumps({'chunk_streams': '$cfg', 'value': round(l['value'],1), 'e2e': round(l['e2e']['value'],1), 'unpipelined_ms': round(l['device_ms_per_step_unpipelined'],1)}))" | tee -a gpurun_out/ctx_sweep.jsonl
done
