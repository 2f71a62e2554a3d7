#!/bin/bash
set -u
export ACB_BUILD_TIMELINE=1   # keep the build digest of the instrumented .so valid if anything calls build() on the box
mkdir -p gpurun_out
echo "== timeline KV=1"; ACB_LM_TIMING=1 timeout 300 python profiles/perf_lm_step.py --one 0 --reps 3 > gpurun_out/v5_timeline_kv1.log 2>&1; tail -11 gpurun_out/v5_timeline_kv1.log
echo "== timeline KV=751"; ACB_LM_TIMING=1 timeout 300 python profiles/perf_lm_step.py --one 750 --reps 3 > gpurun_out/v5_timeline_kv751.log 2>&1; tail -11 gpurun_out/v5_timeline_kv751.log
echo "== timeline KV=1 no PDL"; ACB_NO_PDL=1 ACB_LM_TIMING=1 timeout 300 python profiles/perf_lm_step.py --one 0 --reps 3 > gpurun_out/v5_timeline_kv1_nopdl.log 2>&1; tail -11 gpurun_out/v5_timeline_kv1_nopdl.log
