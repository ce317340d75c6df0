#!/bin/bash
# Round-2 opener: parity of the opt-in variants, then the headline bench A/B (one line per setting).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_variants.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MB200_TEST_VARIANTS=1 timeout 900 python -m pytest tests/test_zz_optin_variants.py -m gpu -x -q > gpurun_out/variants_pytest.log 2>&1
tail -3 gpurun_out/variants_pytest.log
: > gpurun_out/variants_ab.jsonl
for cfg in "0 0" "1 0" "2 0" "3 0" "0 1" "0 2" "0 5" "0 7" "0 six" "1 six" "0 g2" "0 g1" "0 g12" "1 g2" "1 5" "0 0" "1 0" "3 0" "0 1" "0 5" "0 six" "0 g2" "0 g1" "0 g12"; do
  set -- $cfg; ls=$1; ns=$2; six=0; g2=0
  g1=0
  if [ "$ns" = "g2" ]; then ns=0; g2=1; fi
  if [ "$ns" = "g1" ]; then ns=0; g1=1; fi
  if [ "$ns" = "g12" ]; then ns=0; g1=1; g2=1; fi
  if [ "$ns" = "six" ]; then ns=0; six=1; fi
  MB200_ACC_LOCKSTEP=$ls MB200_NTT_SMEM=$ns MB200_H_SIX=$six MB200_ACC_G2_SMEM=$g2 MB200_ACC_G1_SMEM=$g1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-circuit-path 2>/dev/null | \
    python -c "import sys,json; l=json.loads(sys.stdin.readline()); print(json.dumps({'acc_lockstep': $ls, 'ntt_smem': $ns, 'h_six': $six, 'g2_smem': $g2, 'g1_smem': $g1, 'value': round(l['value'],1), 'e2e': round(l['e2e']['value'],1), 'unpipelined_ms': round(l['device_ms_per_step_unpipelined'],1), 'acc_share': round(l['roofline']['share_of_step'],3), 'sm_mhz': l['clocks']['sm_mhz']}))" | tee -a gpurun_out/variants_ab.jsonl
done
# launch list of the lock-step kernel alone (cold, serialised): per-launch time against msm_accumulate_g1
MB200_ACC_LOCKSTEP=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__pcsamp_warps_issue_stalled_no_instructions,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:msm_accumulate_g1 -c 6 --csv --log-file gpurun_out/variants_ncu_lockstep.csv \
  python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-circuit-path > /dev/null 2>&1
tail -8 gpurun_out/variants_ncu_lockstep.csv
MB200_NTT_SMEM=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:ntt_fused -c 8 --csv --log-file gpurun_out/variants_ncu_ntt_fused.csv \
  python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-circuit-path > /dev/null 2>&1
tail -10 gpurun_out/variants_ncu_ntt_fused.csv
