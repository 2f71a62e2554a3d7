#!/bin/bash
# Round 2, GPU session 11: MMA issuer + barrier poller off the producer warp scheduler; cheap tile cursor.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
for NCW in 8; do
  if [ "$NCW" != "8" ]; then ACB_STEP_NCW=$NCW python -m audiocraft_b200.build --force > gpurun_out/r2s11_build_$NCW.log 2>&1; fi
  echo "== NCW=$NCW phases lm_mini"; $T 120 python tests/debug_fused.py phases lm_mini > gpurun_out/r2s11_phases_mini_$NCW.log 2>&1; echo "rc=$?"; grep -E "BAD|PHASES" gpurun_out/r2s11_phases_mini_$NCW.log | head -8
  echo "== NCW=$NCW e2e medium_2l"; $T 180 python tests/debug_fused.py e2e lm_medium_2l 8 > gpurun_out/r2s11_e2e_m2l_$NCW.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s11_e2e_m2l_$NCW.log
  echo "== NCW=$NCW e2e large_2l rows 64"; $T 180 python tests/debug_fused.py e2e lm_large_2l 32 > gpurun_out/r2s11_e2e_l2l_$NCW.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s11_e2e_l2l_$NCW.log
  echo "== NCW=$NCW trace"; ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 0 --reps 2 > gpurun_out/r2s11_trace_kv1_$NCW.log 2>&1; echo "rc=$?"; grep -A 8 "step trace" gpurun_out/r2s11_trace_kv1_$NCW.log | tail -9
  ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 1499 --reps 2 > gpurun_out/r2s11_trace_kv1500_$NCW.log 2>&1; grep "step trace\] rows" gpurun_out/r2s11_trace_kv1500_$NCW.log | tail -1
  echo "== NCW=$NCW perf"; $T 300 python profiles/perf_lm_step.py > gpurun_out/r2s11_perf_fused_$NCW.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s11_perf_fused_$NCW.log
done
