#!/bin/bash
# Round 2, GPU session 14: first hardware contact of the cluster split-K / LayerNorm-on-load step (v10) -- bring-up vs the v9 layer,
# step timing at pinned KV lengths for both, LM test files.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
for a in "lm_mini 2" "lm_mini 5" "lm_medium_2l 8" "lm_medium_2l 20" "lm_large_2l 8"; do
  set -- $a
  echo "== debug_v10 $1 B=$2"; $T 240 python tests/debug_v10.py $1 $2 > gpurun_out/r2s14_debug_$1_$2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s14_debug_$1_$2.log
done
echo "== step perf v10"; ACB_LM_STEP=v10 $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s14_perf_step_v10.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s14_perf_step_v10.log
echo "== step perf v9"; ACB_LM_STEP=v9 $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s14_perf_step_v9.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s14_perf_step_v9.log
echo "== LM tests"; $T 900 python -m pytest tests/test_gpu_lm.py -q -m gpu -x > gpurun_out/r2s14_pytest_lm.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s14_pytest_lm.log
