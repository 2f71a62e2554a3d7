#!/bin/bash
# Round 2, GPU session 19: v10 step with TMA-staged activation slices (fp32 slice normalised in place), compact code.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== debug_v10 medium_2l"; $T 240 python tests/debug_v10.py lm_medium_2l 8 > gpurun_out/r2s19_debug_v10.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s19_debug_v10.log
echo "== debug_v10 large_2l B=20 (64 rows)"; $T 240 python tests/debug_v10.py lm_large_2l 20 > gpurun_out/r2s19_debug_v10_large.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s19_debug_v10_large.log
echo "== debug_v10 mini B=3"; $T 240 python tests/debug_v10.py lm_mini 3 > gpurun_out/r2s19_debug_v10_mini.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s19_debug_v10_mini.log
echo "== step perf v10"; ACB_LM_STEP=v10 $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s19_perf_step_v10.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s19_perf_step_v10.log
echo "== timeline v10 KV=1"; ACB_LIB=$PWD/audiocraft_b200/libaudiocraft_b200_timeline.so ACB_LM_STEP=v10 ACB_LM_TIMING=1 $T 300 python profiles/perf_lm_step.py --one 0 --reps 3 > gpurun_out/r2s19_timeline_v10_kv1.log 2>&1; tail -9 gpurun_out/r2s19_timeline_v10_kv1.log
echo "== LM tests"; $T 900 python -m pytest tests/test_gpu_lm.py -q -m gpu -x > gpurun_out/r2s19_pytest_lm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s19_pytest_lm.log
