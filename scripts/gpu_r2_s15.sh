#!/bin/bash
# Round 2, GPU session 15 (instrumented build): layer-0 timeline of the v10 step (cluster GEMMs) and of the v9 step at KV 1.
set -u
export ACB_BUILD_TIMELINE=1
mkdir -p gpurun_out
echo "== timeline v10 KV=1"; ACB_LM_STEP=v10 ACB_LM_TIMING=1 timeout -s KILL 300 python profiles/perf_lm_step.py --one 0 --reps 3 > gpurun_out/r2s15_timeline_v10_kv1.log 2>&1; tail -9 gpurun_out/r2s15_timeline_v10_kv1.log
echo "== timeline v9 KV=1"; ACB_LM_STEP=v9 ACB_LM_TIMING=1 timeout -s KILL 300 python profiles/perf_lm_step.py --one 0 --reps 3 > gpurun_out/r2s15_timeline_v9_kv1.log 2>&1; tail -12 gpurun_out/r2s15_timeline_v9_kv1.log
echo "== v10 without PDL"; ACB_NO_PDL=1 ACB_LM_STEP=v10 timeout -s KILL 300 python profiles/perf_lm_step.py > gpurun_out/r2s15_perf_v10_nopdl.log 2>&1; head -3 gpurun_out/r2s15_perf_v10_nopdl.log
