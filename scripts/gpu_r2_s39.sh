#!/bin/bash
# Round 2, GPU session 39: LSTM step with both operands pre-split into fp16 terms (lstm_h2_kernel) vs the 3xTF32 kernel.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== lstm tests"; $T 300 python -m pytest tests/test_gpu_encodec.py -q -m gpu -s -k "lstm" > gpurun_out/r2s39_pytest_lstm.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s39_pytest_lstm.log
echo "== encodec + fullsize tests"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -s -k "not lm and not medium and not large" > gpurun_out/r2s39_pytest_encodec.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s39_pytest_encodec.log; grep -E "latent max err" gpurun_out/r2s39_pytest_encodec.log | head -6
echo "== encodec perf (fp16x2 LSTM)"; $T 300 python profiles/perf_encodec.py > gpurun_out/r2s39_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s39_perf_encodec.log
echo "== encodec perf (3xTF32 LSTM)"; ACB_LSTM_TC=3 $T 300 python profiles/perf_encodec.py > gpurun_out/r2s39_perf_encodec_tf32.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s39_perf_encodec_tf32.log
