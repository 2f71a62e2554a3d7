#!/bin/bash
# Round 2, GPU session 43: lstm_h2_kernel with all 8 k16 steps of B fragments requested at once (variant build) vs 4 + 4.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== default"; $T 200 python profiles/perf_encodec.py > gpurun_out/r2s43_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s43_perf_encodec.log
echo "== PB=8"; ACB_LIB=$PWD/audiocraft_b200/libaudiocraft_b200_pb8.so $T 200 python profiles/perf_encodec.py > gpurun_out/r2s43_perf_encodec_pb8.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s43_perf_encodec_pb8.log
