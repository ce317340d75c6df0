#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python -c "
import masp_b200.prover as pv
pv.init(0)
print('fpmul/s %.3e' % pv.bench_fpmul())
for m,n in ((0,'mul inline'),(1,'mul out-of-line'),(2,'xyzz dbl'),(3,'xyzz add')): print('%-16s %.0f ns/op' % (n, pv.bench_latency(m)))
" 2>&1 | tee gpurun_out/latency.log
echo "== msm sweep"; timeout 900 python scripts/msm_sweep.py --sizes 16 18 20 22 --check 2>&1 | tee gpurun_out/msm_sweep_n1.log
echo "== mixed"; timeout 600 python scripts/mixed_batch.py --mode mixed --tx 32 --check 1 2>&1 | tee gpurun_out/mixed_n1.log
echo "== convert"; timeout 600 python scripts/mixed_batch.py --mode convert --per-gpu 128 --check 2 2>&1 | tee gpurun_out/convert_n1.log
echo "== single proof latency"; timeout 300 python -c "
import time, masp_b200.prover as pv
from masp_b200 import synthetic as syn
pv.init(0)
for name in ('output','convert','spend'):
    sh = syn.SHAPES[name]
    P = pv.Parameters.read(pv.params_synthesize(sh), sh.densities())
    w = syn.witness(sh, 0, pv.fr_mul)
    a = pv.ProvingAssignment(w['a'], w['b'], w['c'], w['inputs'], w['aux'])
    pv.create_proof(a, P, w['r'], w['s'])
    t0=time.perf_counter()
    for _ in range(5): pv.create_proof(a, P, w['r'], w['s'])
    print(name, 'single-proof latency %.1f ms' % ((time.perf_counter()-t0)/5*1e3))
" 2>&1 | tee gpurun_out/single_latency.log
