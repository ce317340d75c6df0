#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for n in 2 4 $N; do
echo "== msm sweep N=$n"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n scripts/msm_sweep.py --sizes 20 22 24 --kinds U --check > gpurun_out/msm_sweep_n$n.log 2>&1; echo "exit $?"; grep '^{' gpurun_out/msm_sweep_n$n.log | cut -c1-330
done
