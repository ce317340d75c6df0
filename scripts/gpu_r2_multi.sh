#!/bin/bash
# Multi-GPU checks: gpurun --gpus N -- 'bash scripts/gpu_r2_multi.sh N'
#  - tests/test_multi_device.py on real devices (one process, N GPUs behind the C ABI)
#  - bench.py under torchrun (the driver's launch) and as ONE process (--gpus N without torchrun)
#  - the MSM sweep both ways (NCCL all-gather of device partials / peer copies inside the library)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
what=${2:-all}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$what" = all ] || [ "$what" = tests ]; then
  timeout 600 python -m pytest tests/test_multi_device.py -m gpu -q 2>&1 | tail -3
fi
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("N=%d value %.1f e2e %.1f ms/step %.1f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]))
        for k, v in d.get("configs", {}).items():
            if k != "msm_sweep": print("  ", k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a in ("proofs_per_s", "tx_per_s", "seconds", "oracle_sample")})
            else:
                for r in v: print("   msm 2^%d %s total %.1f ms acc %.1f ms ok=%s" % (r["log_n"], r["scalars"], r["ms_total"], r["ms_accumulate_kernel"], r["closed_form_ok"]))
        cp = d.get("circuit_path") or {}
        print("   circuit_path", {a: (round(b, 1) if isinstance(b, float) else b) for a, b in cp.items() if a in ("proofs_per_s_pipelined", "host_witness_per_s", "vs_synthetic_rows_e2e", "error")})
        print("   cpu", (d.get("cpu_baseline") or {}).get("gpu_proofs_byte_identical"), d.get("parity_error"))'
if [ "$what" = all ] || [ "$what" = torchrun ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 \
     > gpurun_out/r02_bench_n${N}_torchrun.json 2> gpurun_out/r02_bench_n${N}_torchrun.err
  python -c "$P" < gpurun_out/r02_bench_n${N}_torchrun.json; tail -3 gpurun_out/r02_bench_n${N}_torchrun.err
fi
if [ "$what" = all ] || [ "$what" = single ]; then
  timeout 900 python bench.py --gpus $N --steps 4 --warmup 3 --msm-sizes 16 20 24 > gpurun_out/r02_bench_n${N}_oneprocess.json 2> gpurun_out/r02_bench_n${N}_oneprocess.err
  python -c "$P" < gpurun_out/r02_bench_n${N}_oneprocess.json; tail -3 gpurun_out/r02_bench_n${N}_oneprocess.err
fi
if [ "$what" = all ] || [ "$what" = msm ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/msm_sweep.py --sizes 16 18 20 22 24 --reps 3 \
     > gpurun_out/r02_msm_sweep_n${N}.jsonl 2> gpurun_out/r02_msm_sweep_n${N}.err
  tail -2 gpurun_out/r02_msm_sweep_n${N}.err
  timeout 600 python scripts/msm_sweep.py --devices $N --sizes 16 20 24 --reps 3 > gpurun_out/r02_msm_sweep_n${N}_oneprocess.jsonl 2> gpurun_out/r02_msm_sweep_n${N}_oneprocess.err
  python - <<PY
import json
for f in ("gpurun_out/r02_msm_sweep_n${N}.jsonl", "gpurun_out/r02_msm_sweep_n${N}_oneprocess.jsonl"):
    print(f)
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print("  2^%d %s total %.1f acc %.1f limit %.1f %s ok=%s" % (d["log_n"], d["scalars"], d["ms_total"], d["ms_accumulate_kernel"], d["limit_ms_1.3x_acc_plus_5"], d["within_limit"], d["closed_form_ok"]))
PY
fi
