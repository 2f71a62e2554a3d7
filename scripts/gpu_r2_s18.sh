#!/bin/bash
# Round 2, GPU session 18: v10 step with the DSMEM push + mbarrier split-K hand-off and the single-pass statistics merge;
# acb_resblock with batched / prefetched staging.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== debug_v10 medium_2l"; $T 240 python tests/debug_v10.py lm_medium_2l 8 > gpurun_out/r2s18_debug_v10.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s18_debug_v10.log
echo "== debug_v10 large_2l B=20 (64 rows)"; $T 240 python tests/debug_v10.py lm_large_2l 20 > gpurun_out/r2s18_debug_v10_large.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s18_debug_v10_large.log
echo "== step perf v10"; ACB_LM_STEP=v10 $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s18_perf_step_v10.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s18_perf_step_v10.log
echo "== resblock tests"; $T 300 python -m pytest tests/test_gpu_encodec.py -q -m gpu -s -k "resblock" > gpurun_out/r2s18_pytest_resblock.log 2>&1; echo "rc=$?"; grep -E "max err|passed|failed|Error" gpurun_out/r2s18_pytest_resblock.log | tail -20
echo "== encodec tests"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -k "not resblock and not lm and not medium and not large" > gpurun_out/r2s18_pytest_encodec.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s18_pytest_encodec.log
echo "== encodec perf"; $T 400 python profiles/perf_encodec.py > gpurun_out/r2s18_perf_encodec.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2s18_perf_encodec.log
echo "== LM tests"; $T 900 python -m pytest tests/test_gpu_lm.py -q -m gpu -x > gpurun_out/r2s18_pytest_lm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s18_pytest_lm.log
