#!/bin/bash
# Round 2, GPU session 17 (instrumented build): layer-0 timeline of the v10 step after the load batching.
set -u
export ACB_BUILD_TIMELINE=1
mkdir -p gpurun_out
echo "== timeline v10 KV=1"; ACB_LM_STEP=v10 ACB_LM_TIMING=1 timeout -s KILL 300 python profiles/perf_lm_step.py --one 0 --reps 3 > gpurun_out/r2s17_timeline_v10_kv1.log 2>&1; tail -9 gpurun_out/r2s17_timeline_v10_kv1.log
