#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in "32 2" "64 2" "64 3" "128 2"; do set -- $cfg
  echo "== chunk $1 streams $2"; timeout 300 python bench.py --batch 256 --steps 2 --warmup 1 --chunk $1 --streams $2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f e2e %.1f ms %.1f share %.3f acc_ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['roofline']['avg_launch_ms']))
    else: print(l, end='')"
done
echo "== batch 1024 chunk 64"; timeout 300 python bench.py --batch 1024 --steps 1 --warmup 1 --chunk 64 --no-cpu-baseline 2>&1 | tail -c 400
