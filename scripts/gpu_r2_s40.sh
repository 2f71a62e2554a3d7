#!/bin/bash
# Round 2, GPU session 40: residual block with operands split into fp16 terms (resblock_h2_kernel) vs the 3xTF32 kernel.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== resblock tests"; $T 300 python -m pytest tests/test_gpu_encodec.py -q -m gpu -s -k "resblock" > gpurun_out/r2s40_pytest_resblock.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s40_pytest_resblock.log; grep -E "fp16x2.*exact=1" gpurun_out/r2s40_pytest_resblock.log | head -8
echo "== encodec + fullsize tests"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -s -k "not lm and not medium and not large and not resblock" > gpurun_out/r2s40_pytest_encodec.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s40_pytest_encodec.log; grep -E "latent max err" gpurun_out/r2s40_pytest_encodec.log | head -6
echo "== encodec perf (fp16-split blocks)"; $T 300 python profiles/perf_encodec.py > gpurun_out/r2s40_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "fused block|layers total" gpurun_out/r2s40_perf_encodec.log
