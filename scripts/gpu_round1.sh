#!/bin/bash
# First GPU pass: smoke, parity tests, a short bench.  Everything under timeouts.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== fpmul"; timeout 120 python -c "
import masp_b200.prover as pv
pv.init(0); print('selftest', pv.selftest()); print('fpmul/s %.3e' % pv.bench_fpmul())" > gpurun_out/fpmul.log 2>&1; cat gpurun_out/fpmul.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 600 python bench.py --batch 32 --steps 2 --warmup 1 --cpu-seconds 4 > gpurun_out/bench_small.log 2>&1; echo "bench exit $?"; tail -c 3000 gpurun_out/bench_small.log
