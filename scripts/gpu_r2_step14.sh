#!/bin/bash
# Round 2, fourteenth GPU call: proof_cmul with inlined doublings, inlined reductions for tiny batches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "prove or tx_prover or extreme or witness_only" 2>&1 | tail -2
timeout 200 python scripts/single_proof.py > gpurun_out/r02_single_proof_latency_b.txt 2>&1; cat gpurun_out/r02_single_proof_latency_b.txt
timeout 300 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/r02_ab_cmul.json 2> gpurun_out/r02_ab_cmul.err
python -c 'import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))' < gpurun_out/r02_ab_cmul.json; tail -2 gpurun_out/r02_ab_cmul.err
