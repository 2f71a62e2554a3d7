#!/bin/bash
# Round 2, GPU session 31: U = 6 activation-load batches in the step GEMM (variant build) vs the default; bench with the cp.async attention.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== step perf default"; $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s31_perf_step_default.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s31_perf_step_default.log
echo "== step perf U=6"; ACB_LIB=$PWD/audiocraft_b200/libaudiocraft_b200_u6.so $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s31_perf_step_u6.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s31_perf_step_u6.log
echo "== bench (no reference-GPU pass)"; $T 600 python bench.py --steps 1 --warmup 3 --no-ref-gpu > gpurun_out/r2s31_bench.json 2> gpurun_out/r2s31_bench.err; echo "rc=$?"; cut -c1-900 gpurun_out/r2s31_bench.json
