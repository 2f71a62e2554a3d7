#!/bin/bash
# Round 2, GPU session 42: ncu full-set captures of the kernels that became defaults at the end of the round (resblock_h2, lstm_h2,
# lm_attn2 at KV 751).
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
for k in resblock_h2 lstm_h2; do
  echo "== ncu $k"; $T 200 ncu --set full --clock-control none -k regex:$k -c 1 -f -o gpurun_out/r2_prof5_$k python profiles/perf_encodec.py --batch 8 > gpurun_out/r2s42_ncu_$k.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2s42_ncu_$k.log
  ncu -i gpurun_out/r2_prof5_$k.ncu-rep --page details --csv > gpurun_out/r2_prof5_${k}_details.csv 2>/dev/null; rm -f gpurun_out/r2_prof5_$k.ncu-rep
done
echo "== ncu lm_attn2"; $T 200 ncu --set full --clock-control none -k regex:lm_attn2 -s 48 -c 1 -f -o gpurun_out/r2_prof5_lm_attn2 python profiles/perf_lm_step.py --one 750 --reps 2 > gpurun_out/r2s42_ncu_attn2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2s42_ncu_attn2.log
ncu -i gpurun_out/r2_prof5_lm_attn2.ncu-rep --page details --csv > gpurun_out/r2_prof5_lm_attn2_details.csv 2>/dev/null; rm -f gpurun_out/r2_prof5_lm_attn2.ncu-rep
