#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest lm"; timeout 900 python -m pytest tests/test_gpu_lm.py -x -q -m gpu > gpurun_out/o2_pytest_lm.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/o2_pytest_lm.log
echo "== perf"; timeout 300 python profiles/perf_lm_step.py > gpurun_out/o2_perf.log 2>&1; cat gpurun_out/o2_perf.log
