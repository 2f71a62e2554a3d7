#!/bin/bash
# Round 2, GPU session 24: ncu full set of conv1d_t6 after the staging fix; LSTM (8 warps x 4 k-steps) confirmation.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== lstm + golden tests"; $T 300 python -m pytest tests/test_gpu_encodec.py -q -m gpu -k "lstm or golden" > gpurun_out/r2s24_pytest_lstm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s24_pytest_lstm.log
echo "== encodec perf"; $T 400 python profiles/perf_encodec.py > gpurun_out/r2s24_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s24_perf_encodec.log
for k in conv1d_t6; do
  echo "== ncu $k"; $T 420 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r2_prof_${k}_v2 python profiles/perf_encodec.py --batch 8 > gpurun_out/r2s24_ncu_$k.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s24_ncu_$k.log
  ncu -i gpurun_out/r2_prof_${k}_v2.ncu-rep --page details --csv > gpurun_out/r2_prof_${k}_v2_details.csv 2>/dev/null
  ncu -i gpurun_out/r2_prof_${k}_v2.ncu-rep --page source --csv > gpurun_out/r2_prof_${k}_v2_source.csv 2>/dev/null
done
