#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== bench full (batch 256)"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "exit $?"; tail -c 2500 gpurun_out/bench_full.log
for cfg in "16 2" "32 2" "64 2" "32 3" "64 1"; do set -- $cfg
  echo "== chunk $1 streams $2"; timeout 300 python bench.py --batch 128 --steps 2 --warmup 1 --chunk $1 --streams $2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f e2e %.1f ms %.1f share %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['share_of_step']))
    else: print(l, end='')"
done
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1; echo "exit $?"
echo "== ncu full g1 accumulate"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_g1 -s 2 -c 2 -o gpurun_out/prof_acc_g1 python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1; echo "exit $?"
ls -la gpurun_out
