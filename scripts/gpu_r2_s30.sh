#!/bin/bash
# Round 2, GPU session 30: conv1d_t6 with the cp.async raw-sample ring; self-attention with the cp.async K/V ring (lm_attn2_kernel).
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== encodec tests"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -k "not lm and not medium and not large" > gpurun_out/r2s30_pytest_encodec.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s30_pytest_encodec.log
echo "== encodec perf"; $T 400 python profiles/perf_encodec.py > gpurun_out/r2s30_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "encoder.model.(3|6|9|12|10|15)|layers total" gpurun_out/r2s30_perf_encodec.log
echo "== LM tests"; $T 900 python -m pytest tests/test_gpu_lm.py -q -m gpu -x > gpurun_out/r2s30_pytest_lm.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s30_pytest_lm.log
echo "== step perf attn2"; $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s30_perf_step_attn2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s30_perf_step_attn2.log
echo "== step perf attn v1"; ACB_LM_ATTN=v1 $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s30_perf_step_attn1.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s30_perf_step_attn1.log
