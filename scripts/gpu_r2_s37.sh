#!/bin/bash
# Round 2, GPU session 37: ncu full set (source-level stall samples) of the four heaviest codec kernels on the final tree.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
for k in conv1d_t5 resblock lstm_tc conv1d_t6; do
  echo "== ncu $k"; $T 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r2_prof4_$k python profiles/perf_encodec.py --batch 8 > gpurun_out/r2s37_ncu_$k.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2s37_ncu_$k.log
  ncu -i gpurun_out/r2_prof4_$k.ncu-rep --page details --csv > gpurun_out/r2_prof4_${k}_details.csv 2>/dev/null
  ncu -i gpurun_out/r2_prof4_$k.ncu-rep --page source --csv > gpurun_out/r2_prof4_${k}_source.csv 2>/dev/null
  rm -f gpurun_out/r2_prof4_$k.ncu-rep
done
