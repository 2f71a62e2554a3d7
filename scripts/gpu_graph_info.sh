#!/bin/bash
mkdir -p gpurun_out
for v in v5 v6; do
ACB_LM_STEP=$v ACB_LM_GRAPH_INFO=1 timeout 300 python profiles/perf_lm_step.py --one 10 2>&1 | grep "acb graph" | tee -a gpurun_out/graph_info.log
done
