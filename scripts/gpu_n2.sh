#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
echo "== N=2 bench"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "exit $?"; tail -c 1500 gpurun_out/bench_n2.log
echo "== N=2 reference arm"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "exit $?"; tail -c 800 gpurun_out/bench_ref_n2.log
echo "== N=1 bench default"; timeout 900 python bench.py > gpurun_out/bench_n1.log 2>&1; echo "exit $?"; tail -c 600 gpurun_out/bench_n1.log
