#!/bin/bash
# Round 2, GPU session 28: ncu full set of the persistent conv1d_t6 (model.3 shape), source-level stall samples.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
k=conv1d_t6
echo "== ncu $k"; $T 420 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r2_prof_${k}_v3 python profiles/perf_encodec.py --batch 8 > gpurun_out/r2s28_ncu_$k.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s28_ncu_$k.log
ncu -i gpurun_out/r2_prof_${k}_v3.ncu-rep --page details --csv > gpurun_out/r2_prof_${k}_v3_details.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_${k}_v3.ncu-rep --page source --csv > gpurun_out/r2_prof_${k}_v3_source.csv 2>/dev/null
