#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== in-kernel stamps, PDL on"; ACB_LM_TIMING=1 timeout 300 python profiles/perf_lm_step.py --one 750 --reps 4 > gpurun_out/v6_timing_pdl.log 2>&1; tail -7 gpurun_out/v6_timing_pdl.log
echo "== in-kernel stamps, PDL off"; ACB_NO_PDL=1 ACB_LM_TIMING=1 timeout 300 python profiles/perf_lm_step.py --one 750 --reps 4 > gpurun_out/v6_timing_nopdl.log 2>&1; tail -7 gpurun_out/v6_timing_nopdl.log
