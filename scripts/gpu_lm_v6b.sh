#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== ncu launch list v6 (one direct step at KV 750)"
timeout 600 ncu -k regex:lm_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/v6_launches_step750.csv python profiles/perf_lm_step.py --one 750 > gpurun_out/v6_ncu.log 2>&1; echo "ncu rc=$?"
echo "== perf v6 no PDL"; ACB_NO_PDL=1 timeout 300 python profiles/perf_lm_step.py > gpurun_out/v6_perf_nopdl.log 2>&1; cat gpurun_out/v6_perf_nopdl.log
echo "== perf v6 slab=100 fill=0 (clusters only for FFN2)"; ACB_LM_SLAB_KB=100 ACB_LM_FILL=0 timeout 300 python profiles/perf_lm_step.py > gpurun_out/v6_perf_s100_f0.log 2>&1; cat gpurun_out/v6_perf_s100_f0.log
