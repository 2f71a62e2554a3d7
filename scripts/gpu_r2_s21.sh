#!/bin/bash
# Round 2, GPU session 21: tensor-core LSTM step (lstm_tc_kernel) -- oracle tests, whole-model goldens, per-layer timing.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== lstm tests"; $T 300 python -m pytest tests/test_gpu_encodec.py -q -m gpu -s -k "lstm" > gpurun_out/r2s21_pytest_lstm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s21_pytest_lstm.log
echo "== encodec tests"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -k "not lm and not medium and not large" > gpurun_out/r2s21_pytest_encodec.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s21_pytest_encodec.log
echo "== encodec perf"; $T 400 python profiles/perf_encodec.py > gpurun_out/r2s21_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s21_perf_encodec.log
echo "== encodec perf, FMA lstm"; ACB_LSTM_TC=0 $T 400 python profiles/perf_encodec.py > gpurun_out/r2s21_perf_encodec_fma_lstm.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s21_perf_encodec_fma_lstm.log
