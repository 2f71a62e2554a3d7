#!/bin/bash
# Round 2, GPU session 4: compact (code-size-aware) fused step.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== gen lm_mini"; $T 120 python tests/debug_fused.py gen lm_mini > gpurun_out/r2s4_gen_mini.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s4_gen_mini.log
echo "== e2e medium_2l"; $T 180 python tests/debug_fused.py e2e lm_medium_2l 8 > gpurun_out/r2s4_e2e_m2l.log 2>&1; echo "rc=$?"; tail -10 gpurun_out/r2s4_e2e_m2l.log
echo "== LM tests on the fused step"; $T 600 python -m pytest tests/test_gpu_lm.py -q -m gpu -k "not wide and not chain and not split_kv and not ft32" > gpurun_out/r2s4_pytest_lm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2s4_pytest_lm.log
echo "== trace"; ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 0 --reps 2 > gpurun_out/r2s4_trace_kv1.log 2>&1; echo "rc=$?"; grep -A 8 "step trace" gpurun_out/r2s4_trace_kv1.log | tail -18
ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 1499 --reps 2 > gpurun_out/r2s4_trace_kv1500.log 2>&1; grep "step trace\] rows" gpurun_out/r2s4_trace_kv1500.log | tail -1
echo "== perf"; $T 300 python profiles/perf_lm_step.py > gpurun_out/r2s4_perf_fused.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s4_perf_fused.log
