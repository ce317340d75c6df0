#!/bin/bash
# Round 2, thirteenth GPU call: binary-Euclid inversion -- self-test / verifier / prover parity, single-proof
# latency, standalone MSM tails, one bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q -k "selftest or verify or prove or msm or tx_prover or sapling" 2>&1 | tail -2
timeout 200 python scripts/single_proof.py > gpurun_out/r02_single_proof_latency.txt 2>&1; cat gpurun_out/r02_single_proof_latency.txt
timeout 300 python scripts/msm_sweep.py --sizes 16 18 20 22 --reps 3 > gpurun_out/r02_msm_sweep_n1_step13.jsonl 2> gpurun_out/r02_msm_sweep_n1_step13.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_msm_sweep_n1_step13.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["log_n"], d["scalars"], "total %.1f acc %.1f limit %.1f %s ok=%s" % (d["ms_total"], d["ms_accumulate_kernel"], d["limit_ms_1.3x_acc_plus_5"], d["within_limit"], d["closed_form_ok"]))
PY
timeout 300 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/r02_ab_gcd.json 2> gpurun_out/r02_ab_gcd.err
python -c 'import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))' < gpurun_out/r02_ab_gcd.json; tail -2 gpurun_out/r02_ab_gcd.err
