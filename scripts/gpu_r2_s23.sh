#!/bin/bash
# Round 2, GPU session 23: conv1d_t6 with batched staging loads, 16-warp lstm_tc; EnCodec tests + per-layer timing.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== encodec tests"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -k "not lm and not medium and not large" > gpurun_out/r2s23_pytest_encodec.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s23_pytest_encodec.log
echo "== encodec perf"; $T 400 python profiles/perf_encodec.py > gpurun_out/r2s23_perf_encodec.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2s23_perf_encodec.log
