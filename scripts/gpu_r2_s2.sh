#!/bin/bash
# Round 2, GPU session 2: bring-up of the persistent fused decode step (every stage under its own short timeout).
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== phases lm_mini"; $T 120 python tests/debug_fused.py phases lm_mini > gpurun_out/r2s2_phases_mini.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/r2s2_phases_mini.log
echo "== e2e lm_mini"; $T 120 python tests/debug_fused.py e2e lm_mini > gpurun_out/r2s2_e2e_mini.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2s2_e2e_mini.log
echo "== gen lm_mini"; $T 120 python tests/debug_fused.py gen lm_mini > gpurun_out/r2s2_gen_mini.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s2_gen_mini.log
echo "== phases medium_2l rows 16"; $T 180 python tests/debug_fused.py phases lm_medium_2l 8 > gpurun_out/r2s2_phases_m2l.log 2>&1; echo "rc=$?"; grep -E "BAD|PHASES|plan" gpurun_out/r2s2_phases_m2l.log | head -30
echo "== e2e medium_2l"; $T 180 python tests/debug_fused.py e2e lm_medium_2l 8 > gpurun_out/r2s2_e2e_m2l.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s2_e2e_m2l.log
echo "== e2e large_2l rows 64"; $T 180 python tests/debug_fused.py e2e lm_large_2l 32 > gpurun_out/r2s2_e2e_l2l.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s2_e2e_l2l.log
echo "== LM test-suite on the fused step"; $T 600 python -m pytest tests/test_gpu_lm.py -x -q -m gpu > gpurun_out/r2s2_pytest_lm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2s2_pytest_lm.log
echo "== perf fused"; ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 0 --reps 2 > gpurun_out/r2s2_trace_kv1.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s2_trace_kv1.log
ACB_LM_STEP_TRACE=1 $T 300 python profiles/perf_lm_step.py --one 1499 --reps 2 > gpurun_out/r2s2_trace_kv1500.log 2>&1; tail -4 gpurun_out/r2s2_trace_kv1500.log
$T 300 python profiles/perf_lm_step.py > gpurun_out/r2s2_perf_fused.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s2_perf_fused.log
ACB_LM_STEP=v5 $T 300 python profiles/perf_lm_step.py > gpurun_out/r2s2_perf_v5.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s2_perf_v5.log
