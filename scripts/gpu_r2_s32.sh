#!/bin/bash
# Round 2, GPU session 32: LSTM with two alternating 16-item halves (lstm_tc2_kernel) vs the one-batch tensor-core kernel.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== lstm + model tests"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -k "not lm and not medium and not large" > gpurun_out/r2s32_pytest_encodec.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s32_pytest_encodec.log
echo "== encodec perf (two halves)"; $T 400 python profiles/perf_encodec.py > gpurun_out/r2s32_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s32_perf_encodec.log
echo "== encodec perf (one batch)"; ACB_LSTM_TC=1 $T 400 python profiles/perf_encodec.py > gpurun_out/r2s32_perf_encodec_onebatch.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s32_perf_encodec_onebatch.log
