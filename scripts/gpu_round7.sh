#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in "48 3" "48 4" "64 4" "86 3" "32 4"; do set -- $cfg
  echo "== chunk $1 streams $2"; timeout 300 python bench.py --batch 256 --steps 2 --warmup 1 --chunk $1 --streams $2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f e2e %.1f ms %.1f share %.3f acc_ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['roofline']['avg_launch_ms']))
    else: print(l, end='')"
done
echo "== single convert proof launch list"
cat > /tmp/single.py <<'PY'
import time, sys
import masp_b200.prover as pv
from masp_b200 import synthetic as syn
pv.init(0)
sh = syn.SHAPES[sys.argv[1]]
P = pv.Parameters.read(pv.params_synthesize(sh), sh.densities())
w = syn.witness(sh, 0, pv.fr_mul)
a = pv.ProvingAssignment(w['a'], w['b'], w['c'], w['inputs'], w['aux'])
for _ in range(2):
    t0=time.perf_counter(); pv.create_proof(a, P, w['r'], w['s']); print(sys.argv[1], 'latency %.1f ms' % ((time.perf_counter()-t0)*1e3), 'device %.1f ms' % (pv.get_counter('last_batch_us')/1e3))
PY
timeout 300 python /tmp/single.py convert; timeout 300 python /tmp/single.py spend
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_single_convert.csv python /tmp/single.py convert > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_single_spend.csv python /tmp/single.py spend > /dev/null 2>&1
