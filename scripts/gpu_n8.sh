#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench N=$N"; timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "exit $?"; grep '^{' gpurun_out/bench_n$N.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('n_gpus %d value %.1f e2e %.1f ms %.1f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']))"
tail -3 gpurun_out/bench_n$N.log | cut -c1-300
echo "== msm sweep N=$N"; timeout 900 $TR --master-port 29522 scripts/msm_sweep.py --sizes 20 22 24 --check > gpurun_out/msm_sweep_n$N.log 2>&1; echo "exit $?"; grep '^{' gpurun_out/msm_sweep_n$N.log | cut -c1-330
echo "== mixed N=$N (config 4 shape: 2 Spend + 2 Output + 1 Convert per tx)"; timeout 900 $TR --master-port 29523 scripts/mixed_batch.py --mode mixed --tx 64 --check 1 > gpurun_out/mixed_n$N.log 2>&1; echo "exit $?"; grep '^{' gpurun_out/mixed_n$N.log
echo "== convert N=$N (config 2: 128 per GPU)"; timeout 900 $TR --master-port 29524 scripts/mixed_batch.py --mode convert --per-gpu 128 --check 1 > gpurun_out/convert_n$N.log 2>&1; echo "exit $?"; grep '^{' gpurun_out/convert_n$N.log
