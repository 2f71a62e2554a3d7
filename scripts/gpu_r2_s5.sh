#!/bin/bash
# Round 2, GPU session 5: fused step without spinning waiters (scheduler starvation fix); 16 vs 8 compute warps.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
for NCW in 16 8; do
  if [ "$NCW" != "16" ]; then ACB_STEP_NCW=$NCW python -m audiocraft_b200.build --force > gpurun_out/r2s5_build_$NCW.log 2>&1; fi
  echo "== NCW=$NCW gen lm_mini"; $T 120 python tests/debug_fused.py gen lm_mini > gpurun_out/r2s5_gen_mini_$NCW.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s5_gen_mini_$NCW.log
  echo "== NCW=$NCW e2e medium_2l"; $T 180 python tests/debug_fused.py e2e lm_medium_2l 8 > gpurun_out/r2s5_e2e_m2l_$NCW.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s5_e2e_m2l_$NCW.log
  echo "== NCW=$NCW trace"; ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 0 --reps 2 > gpurun_out/r2s5_trace_kv1_$NCW.log 2>&1; echo "rc=$?"; grep -A 8 "step trace" gpurun_out/r2s5_trace_kv1_$NCW.log | tail -9
  ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 1499 --reps 2 > gpurun_out/r2s5_trace_kv1500_$NCW.log 2>&1; grep "step trace\] rows" gpurun_out/r2s5_trace_kv1500_$NCW.log | tail -1
  echo "== NCW=$NCW perf"; $T 300 python profiles/perf_lm_step.py > gpurun_out/r2s5_perf_fused_$NCW.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s5_perf_fused_$NCW.log
done
