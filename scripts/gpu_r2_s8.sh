#!/bin/bash
# Round 2, GPU session 8: L2 weight prefetch cursor, grouped self-attention, wider load rounds.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
for NCW in 8; do
  if [ "$NCW" != "8" ]; then ACB_STEP_NCW=$NCW python -m audiocraft_b200.build --force > gpurun_out/r2s8_build_$NCW.log 2>&1; fi
  echo "== NCW=$NCW phases lm_mini"; $T 120 python tests/debug_fused.py phases lm_mini > gpurun_out/r2s8_phases_mini_$NCW.log 2>&1; echo "rc=$?"; grep -E "BAD|PHASES" gpurun_out/r2s8_phases_mini_$NCW.log | head -8
  echo "== NCW=$NCW e2e medium_2l"; $T 180 python tests/debug_fused.py e2e lm_medium_2l 8 > gpurun_out/r2s8_e2e_m2l_$NCW.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s8_e2e_m2l_$NCW.log
  echo "== NCW=$NCW e2e large_2l rows 64"; $T 180 python tests/debug_fused.py e2e lm_large_2l 32 > gpurun_out/r2s8_e2e_l2l_$NCW.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s8_e2e_l2l_$NCW.log
  echo "== NCW=$NCW trace"; ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 0 --reps 2 > gpurun_out/r2s8_trace_kv1_$NCW.log 2>&1; echo "rc=$?"; grep -A 8 "step trace" gpurun_out/r2s8_trace_kv1_$NCW.log | tail -9
  ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 1499 --reps 2 > gpurun_out/r2s8_trace_kv1500_$NCW.log 2>&1; grep "step trace\] rows" gpurun_out/r2s8_trace_kv1500_$NCW.log | tail -1
  echo "== NCW=$NCW perf"; $T 300 python profiles/perf_lm_step.py > gpurun_out/r2s8_perf_fused_$NCW.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s8_perf_fused_$NCW.log
done
echo "== L2 lookahead sweep"; for LA in 0 12 48; do ACB_LM_L2_AHEAD=$LA timeout -s KILL 200 python profiles/perf_lm_step.py > gpurun_out/r2s8_perf_la$LA.log 2>&1; echo "LA=$LA"; grep "kv_len" gpurun_out/r2s8_perf_la$LA.log | sed -n '1p;5p'; done
echo "== LM tests"; timeout -s KILL 600 python -m pytest tests/test_gpu_lm.py -q -m gpu -k "not wide and not chain and not split_kv and not ft32" > gpurun_out/r2s8_pytest_lm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s8_pytest_lm.log
