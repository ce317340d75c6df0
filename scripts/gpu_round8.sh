#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
P='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.1f e2e %.1f ms %.1f share %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["share_of_step"]))
    else: print(l, end="")'
echo "== bench r8"; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$P"
echo "== bench radix-4 NTT"; MB200_NTT_RADIX_LOG=2 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$P"
echo "== bench streams 4"; timeout 600 python bench.py --steps 4 --warmup 3 --streams 4 --no-cpu-baseline 2>&1 | python -c "$P"
