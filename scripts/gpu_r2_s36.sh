#!/bin/bash
# Round 2, GPU session 36: integer tf32 rounding (bit-identical to cvt.rna for finite inputs) as default; experiment build handing the low
# split term to the tensor core unrounded.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== encodec tests (default build)"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -k "not lm and not medium and not large" > gpurun_out/r2s36_pytest_encodec.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s36_pytest_encodec.log
echo "== encodec perf (default build)"; $T 300 python profiles/perf_encodec.py > gpurun_out/r2s36_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "lstm|fused block|layers total" gpurun_out/r2s36_perf_encodec.log
echo "== lo-raw variant: tests"; ACB_LIB=$PWD/audiocraft_b200/libaudiocraft_b200_lraw.so $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -s -k "not lm and not medium and not large" > gpurun_out/r2s36_pytest_lraw.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s36_pytest_lraw.log; grep -E "resblock C=.*exact=1|latent" gpurun_out/r2s36_pytest_lraw.log | head -12
echo "== lo-raw variant: perf"; ACB_LIB=$PWD/audiocraft_b200/libaudiocraft_b200_lraw.so $T 300 python profiles/perf_encodec.py > gpurun_out/r2s36_perf_encodec_lraw.log 2>&1; echo "rc=$?"; grep -E "lstm|fused block|layers total" gpurun_out/r2s36_perf_encodec_lraw.log
