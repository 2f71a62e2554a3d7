#!/bin/bash
# One GPU session for the v6 LM step: staged bring-up, parity tests, A/B timing against the round-1 kernels, tile-plan sweep.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/v6_gpu.txt 2>&1
echo "== debug bring-up"; timeout 180 python tests/debug_lm.py > gpurun_out/v6_debug.log 2>&1; echo "debug rc=$?"; tail -5 gpurun_out/v6_debug.log
echo "== pytest lm"; timeout 900 python -m pytest tests/test_gpu_lm.py -x -q -m gpu -s > gpurun_out/v6_pytest_lm.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/v6_pytest_lm.log
echo "== perf v6"; timeout 300 python profiles/perf_lm_step.py > gpurun_out/v6_perf.log 2>&1; cat gpurun_out/v6_perf.log
echo "== perf v5"; ACB_LM_STEP=v5 timeout 300 python profiles/perf_lm_step.py > gpurun_out/v6_perf_v5.log 2>&1; cat gpurun_out/v6_perf_v5.log
for cfg in "128 90" "28 90" "56 180"; do
  set -- $cfg
  echo "== perf v6 slab=$1 fill=$2"; ACB_LM_SLAB_KB=$1 ACB_LM_FILL=$2 timeout 300 python profiles/perf_lm_step.py > gpurun_out/v6_perf_s$1_f$2.log 2>&1; cat gpurun_out/v6_perf_s$1_f$2.log
done
echo "== ncu launch list (one direct step at KV 750)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/v6_launches_step750.csv python profiles/perf_lm_step.py --one 750 > gpurun_out/v6_ncu.log 2>&1; echo "ncu rc=$?"
