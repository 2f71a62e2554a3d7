#!/bin/bash
# Round-2 runbook, first GPU contact of the experimental conv1d_t6 kernel (never run on hardware in round 1).
# A tcgen05 / mbarrier bug can hang the GPU: every step is under a short `timeout`, smallest case first.
set -u
mkdir -p gpurun_out
echo "== smallest case"; ACB_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_encodec.py -x -q -s -m gpu \
  -k "experimental_conv1d_t6 and 8-64-7-1-1-False-3" > gpurun_out/t6_first.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t6_first.log
echo "== all unit cases"; ACB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_encodec.py -x -q -s -m gpu \
  -k "experimental_conv1d_t6" > gpurun_out/t6_unit.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t6_unit.log
echo "== model level"; ACB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_encodec.py -x -q -s -m gpu \
  -k "experimental_flush_encoder" > gpurun_out/t6_model.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t6_model.log
echo "== per-layer timing"; timeout 300 python profiles/perf_encodec.py --enc tf32x3_flush > gpurun_out/t6_perf.log 2>&1; tail -40 gpurun_out/t6_perf.log
