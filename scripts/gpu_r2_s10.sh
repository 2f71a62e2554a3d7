#!/bin/bash
# Round 2, GPU session 10: per-K-block stamps inside the MMA loop of the fused step.
set -u
mkdir -p gpurun_out
ACB_LM_STEP_TRACE=1 timeout -s KILL 300 python profiles/perf_lm_step.py --one 0 --reps 2 > gpurun_out/r2s10_trace_kv1.log 2>&1; echo "rc=$?"; grep -A 12 "step trace" gpurun_out/r2s10_trace_kv1.log | tail -13
