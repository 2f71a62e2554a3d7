#!/bin/bash
set -u
export ACB_BUILD_TIMELINE=1   # keep the build digest of the instrumented .so valid if anything calls build() on the box
mkdir -p gpurun_out
echo "== timeline KV=1"; ACB_LM_TIMING=1 timeout 300 python profiles/perf_lm_step.py --one 0 --reps 3 > gpurun_out/t2_timeline_kv1.log 2>&1; tail -11 gpurun_out/t2_timeline_kv1.log
echo "== perf"; timeout 300 python profiles/perf_lm_step.py > gpurun_out/t2_perf.log 2>&1; head -3 gpurun_out/t2_perf.log
