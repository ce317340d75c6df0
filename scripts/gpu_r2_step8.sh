#!/bin/bash
# Round 2, eighth GPU call: G2 accumulate variants as alternate libraries (scripts/debug/abl/lib_g2_*.so):
# (the alternate libraries under scripts/debug/abl were built ad hoc for this one call -- the shipped sources with one macro or constant changed -- and are not kept; the outcomes are in profiles/r02_ab_*.jsonl)
#   A Karatsuba Fp2 (shipped)  B Fp2 mul as two fused two-product reductions  C = B + Y3 as two four-product reductions  D = A + that Y3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp masp_b200/libmasp_b200.so /tmp/lib_orig.so
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f share %.3f acc_ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["roofline"]["avg_launch_ms"]))'
for v in A B C D; do
  cp scripts/debug/abl/lib_g2_$v.so masp_b200/libmasp_b200.so
  timeout 200 python -m pytest tests -m gpu -x -q -k "msm_g2 or prove_tiny" 2>&1 | tail -1
  timeout 400 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/r02_ab_g2_$v.json 2> gpurun_out/r02_ab_g2_$v.err
  echo "variant $v: $(python -c "$P" < gpurun_out/r02_ab_g2_$v.json)"; tail -1 gpurun_out/r02_ab_g2_$v.err
done
cp /tmp/lib_orig.so masp_b200/libmasp_b200.so
