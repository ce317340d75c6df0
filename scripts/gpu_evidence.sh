#!/bin/bash
# Round evidence: tests, bench (N=1), launch list, full ncu capture of the dominant kernel, config drivers.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/bench_n1.log 2>&1; echo "exit $?"; tail -c 700 gpurun_out/bench_n1.log
echo "== single proof"; timeout 300 python scripts/single_proof.py 2>&1 | tee gpurun_out/single_proof.log
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1; echo "exit $?"
echo "== single convert launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_single_convert.csv python scripts/single_proof.py convert > /dev/null 2>&1
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_g1 -s 3 -c 1 -o gpurun_out/prof_acc_g1 python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1; echo "exit $?"
echo "== msm sweep"; timeout 900 python scripts/msm_sweep.py --sizes 16 18 20 22 24 --check 2>&1 | tee gpurun_out/msm_sweep_n1.log | cut -c1-220
echo "== mixed"; timeout 600 python scripts/mixed_batch.py --mode mixed --tx 64 --check 1 2>&1 | tee gpurun_out/mixed_n1.log
echo "== convert"; timeout 600 python scripts/mixed_batch.py --mode convert --per-gpu 128 --check 2 2>&1 | tee gpurun_out/convert_n1.log
