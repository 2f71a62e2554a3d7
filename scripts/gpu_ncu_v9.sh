#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu -k regex:lm_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/v9_launches_step750.csv python profiles/perf_lm_step.py --one 750 > gpurun_out/v9_ncu.log 2>&1; echo "ncu rc=$?"
