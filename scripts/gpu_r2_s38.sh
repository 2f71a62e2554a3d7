#!/bin/bash
# Round 2, GPU session 38: LSTM barrier through red.release / ld.acquire, residual-block slab batches of 34, conv1d_t6 raw ring 4 deep.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== encodec tests"; $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -k "not lm and not medium and not large" > gpurun_out/r2s38_pytest_encodec.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s38_pytest_encodec.log
echo "== encodec perf"; $T 300 python profiles/perf_encodec.py > gpurun_out/r2s38_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "lstm|fused block|encoder.model.(3|6|9|12).conv|layers total" gpurun_out/r2s38_perf_encodec.log
echo "== smoke"; $T 300 python __graft_entry__.py smoke > gpurun_out/r2s38_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s38_smoke.log
