#!/bin/bash
# Round 2, GPU session 22: LSTM step with all B fragments prefetched; ncu captures (full set) of conv1d_t6, resblock, lstm_tc;
# launch list + DRAM bytes of the default LM step at KV 751.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== lstm tests"; $T 300 python -m pytest tests/test_gpu_encodec.py -q -m gpu -k "lstm or golden" > gpurun_out/r2s22_pytest_lstm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s22_pytest_lstm.log
echo "== encodec perf"; $T 400 python profiles/perf_encodec.py > gpurun_out/r2s22_perf_encodec.log 2>&1; echo "rc=$?"; grep -E "lstm|layers total" gpurun_out/r2s22_perf_encodec.log
for k in conv1d_t6 resblock lstm_tc; do
  echo "== ncu $k"; $T 420 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r2_prof_$k python profiles/perf_encodec.py --batch 8 > gpurun_out/r2s22_ncu_$k.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s22_ncu_$k.log
  ncu -i gpurun_out/r2_prof_$k.ncu-rep --page raw --csv > gpurun_out/r2_prof_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_prof_$k.ncu-rep --page details --csv > gpurun_out/r2_prof_${k}_details.csv 2>/dev/null
done
echo "== ncu: launch list + DRAM bytes of the default step at KV 751 (second of two direct steps)"
$T 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:lm_ -s 580 -c 532 --csv --log-file gpurun_out/r2_step_kv751_launches_dram.csv python profiles/perf_lm_step.py --one 750 --reps 2 > gpurun_out/r2s22_ncu_step.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s22_ncu_step.log; wc -l gpurun_out/r2_step_kv751_launches_dram.csv
