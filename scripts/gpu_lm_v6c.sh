#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest lm"; timeout 900 python -m pytest tests/test_gpu_lm.py -x -q -m gpu -s > gpurun_out/v6c_pytest_lm.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v6c_pytest_lm.log
echo "== perf v6c"; timeout 300 python profiles/perf_lm_step.py > gpurun_out/v6c_perf.log 2>&1; cat gpurun_out/v6c_perf.log
echo "== ncu launch list v6c (one direct step at KV 750)"
timeout 600 ncu -k regex:lm_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/v6c_launches_step750.csv python profiles/perf_lm_step.py --one 750 > gpurun_out/v6c_ncu.log 2>&1; echo "ncu rc=$?"
