#!/bin/bash
# Evidence for the second half of round 1: launch list at the final configuration (with the
# real-witness leg and the self-check), full ncu captures of the two new kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --verify --circuit-batch 32 --circuit-rounds 2 > gpurun_out/ncu_launch_final.log 2>&1; echo "exit $?"
python scripts/launch_summary.py gpurun_out/launches_final.csv build_table_g1 build_table_g2 synth_g1 synth_g2 pair_prep decode_g1 decode_g2 > gpurun_out/launch_summary_final.txt; head -30 gpurun_out/launch_summary_final.txt
echo "== ncu full r1cs_eval"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:r1cs_eval -s 1 -c 1 -o gpurun_out/prof_r1cs_eval python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --circuit-batch 64 --circuit-rounds 2 > gpurun_out/ncu_r1cs_run.log 2>&1; echo "exit $?"
echo "== ncu full verify_proofs"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:verify_proofs -s 1 -c 1 -o gpurun_out/prof_verify python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-circuit-path --verify > gpurun_out/ncu_verify_run.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
