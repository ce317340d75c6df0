#!/bin/bash
# Round 2, twelfth GPU call: chunk size / chunk contexts after the kernel changes, then the final test + bench evidence.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f share %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"]))'
: > gpurun_out/r02_chunk_streams_sweep.jsonl
for cs in "64 3" "64 4" "96 3" "128 3" "48 4"; do
  set -- $cs
  timeout 300 python bench.py --steps 4 --warmup 3 --chunk $1 --streams $2 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/tmp_cs.json 2> gpurun_out/tmp_cs.err
  echo "chunk $1 streams $2: $(python -c "$P" < gpurun_out/tmp_cs.json)"
  python - "$1" "$2" <<'PY' >> gpurun_out/r02_chunk_streams_sweep.jsonl
import json, sys
d = json.loads(open("gpurun_out/tmp_cs.json").read().strip().splitlines()[-1])
print(json.dumps({"chunk": int(sys.argv[1]), "streams": int(sys.argv[2]), "value": d["value"], "e2e": d["e2e"]["value"], "ms_per_step": d["ms_per_step"], "clocks": d["clocks"]}))
PY
done
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_final.json").read().strip().splitlines()[-1])
print("FINAL value %.1f e2e %.1f ms/step %.1f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
for k, v in d.get("configs", {}).items():
    if k != "msm_sweep": print(k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a in ("proofs_per_s", "tx_per_s", "oracle_sample")})
print("circuit_path", {a: (round(b, 1) if isinstance(b, float) else b) for a, b in (d.get("circuit_path") or {}).items() if a in ("proofs_per_s_pipelined", "host_witness_per_s", "vs_synthetic_rows_e2e", "error")})
print("cpu", (d.get("cpu_baseline") or {}).get("gpu_proofs_byte_identical"))
PY
