#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f e2e %.1f ms %.1f share %.3f acc_ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['roofline']['avg_launch_ms']))
    else: print(l, end='')"
echo "== convert launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_convert.csv python scripts/mixed_batch.py --mode convert --per-gpu 64 > gpurun_out/ncu_convert_run.log 2>&1; echo "exit $?"
tail -2 gpurun_out/ncu_convert_run.log
