#!/bin/bash
# Round-2 opener for the opt-in kernel variants written (unmeasured) at the end of round 1.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_r2_variants.sh parity'  # byte-exactness of every variant (each child run is cut at 300 s)
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2_variants.sh ab'       # headline bench, one JSON line per setting (~8 min)
#   gpurun --timeout 600 -- 'bash scripts/gpu_r2_variants.sh ncu'      # launch times / stalls of the new kernels
# Settings: "<MB200_ACC_LOCKSTEP> <x>", x = MB200_NTT_SMEM value, or six (MB200_H_SIX), g1 / g2 / g12 (MB200_ACC_G*_SMEM).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
what=${1:-all}

if [ "$what" = parity ] || [ "$what" = all ]; then
  MB200_TEST_VARIANTS=1 timeout 1200 python -m pytest tests/test_zz_optin_variants.py -m gpu -q > gpurun_out/variants_pytest.log 2>&1
  tail -15 gpurun_out/variants_pytest.log
fi

if [ "$what" = ab ] || [ "$what" = all ]; then
  : > gpurun_out/variants_ab.jsonl
  for cfg in "0 0" "1 0" "3 0" "4 0" "0 g1" "0 g2" "0 g12" "0 six" "0 1" "0 2" "0 5" "0 7" "0 0" "1 0" "0 g2" "0 six" "0 2"; do
    set -- $cfg; ls=$1; ns=$2; six=0; g1=0; g2=0
    case "$ns" in six) ns=0; six=1;; g1) ns=0; g1=1;; g2) ns=0; g2=1;; g12) ns=0; g1=1; g2=1;; esac
    MB200_ACC_LOCKSTEP=$ls MB200_NTT_SMEM=$ns MB200_H_SIX=$six MB200_ACC_G1_SMEM=$g1 MB200_ACC_G2_SMEM=$g2 \
      timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-circuit-path 2>/dev/null | \
      python -c "import sys,json; l=json.loads(sys.stdin.readline()); print(json.dumps({'acc_lockstep': $ls, 'ntt_smem': $ns, 'h_six': $six, 'g1_smem': $g1, 'g2_smem': $g2, 'value': round(l['value'],1), 'e2e': round(l['e2e']['value'],1), 'unpipelined_ms': round(l['device_ms_per_step_unpipelined'],1), 'acc_share': round(l['roofline']['share_of_step'],3), 'sm_mhz': l['clocks']['sm_mhz']}))" | tee -a gpurun_out/variants_ab.jsonl
  done
fi

if [ "$what" = ncu ] || [ "$what" = all ]; then
  M=gpu__time_duration.sum,smsp__pcsamp_warps_issue_stalled_no_instructions,smsp__pcsamp_warps_issue_stalled_wait,smsp__pcsamp_warps_issue_stalled_long_scoreboard,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum
  B="python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-circuit-path"
  timeout 500 ncu --metrics $M --clock-control none -k regex:msm_accumulate_g1 -c 6 --csv --log-file gpurun_out/variants_ncu_acc_default.csv $B > /dev/null 2>&1
  MB200_ACC_LOCKSTEP=1 timeout 500 ncu --metrics $M --clock-control none -k regex:msm_accumulate_g1 -c 6 --csv --log-file gpurun_out/variants_ncu_acc_lockstep.csv $B > /dev/null 2>&1
  MB200_ACC_G1_SMEM=1 timeout 500 ncu --metrics $M --clock-control none -k regex:msm_accumulate_g1 -c 6 --csv --log-file gpurun_out/variants_ncu_acc_g1smem.csv $B > /dev/null 2>&1
  MB200_NTT_SMEM=1 timeout 500 ncu --metrics $M --clock-control none -k regex:ntt_fused -c 8 --csv --log-file gpurun_out/variants_ncu_ntt_fused.csv $B > /dev/null 2>&1
  tail -4 gpurun_out/variants_ncu_acc_default.csv gpurun_out/variants_ncu_acc_lockstep.csv gpurun_out/variants_ncu_acc_g1smem.csv gpurun_out/variants_ncu_ntt_fused.csv
fi
