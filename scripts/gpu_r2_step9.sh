#!/bin/bash
# Round 2, ninth GPU call: alternate libraries (scripts/debug/abl): fused NTT on 1024-element blocks (4 per SM),
# (the alternate libraries under scripts/debug/abl were built ad hoc for this one call -- the shipped sources with one macro or constant changed -- and are not kept; the outcomes are in profiles/r02_ab_*.jsonl)
# accumulate segments of 256 / 64 entries; the shipped library first and last.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp masp_b200/libmasp_b200.so /tmp/lib_orig.so
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms/step %.1f share %.3f acc_ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["roofline"]["avg_launch_ms"]))'
for v in orig ntt10 seg256 seg64 orig2; do
  case $v in orig|orig2) cp /tmp/lib_orig.so masp_b200/libmasp_b200.so;; *) cp scripts/debug/abl/lib_$v.so masp_b200/libmasp_b200.so;; esac
  timeout 200 python -m pytest tests -m gpu -x -q -k "ntt or h_coeff or prove_tiny or heavy" 2>&1 | tail -1
  timeout 400 python bench.py --steps 4 --warmup 3 --no-msm-sweep --no-configs --no-cpu-baseline --no-circuit-path > gpurun_out/r02_ab9_$v.json 2> gpurun_out/r02_ab9_$v.err
  echo "variant $v: $(python -c "$P" < gpurun_out/r02_ab9_$v.json)"; tail -1 gpurun_out/r02_ab9_$v.err
done
cp /tmp/lib_orig.so masp_b200/libmasp_b200.so
