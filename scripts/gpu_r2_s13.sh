#!/bin/bash
# Round 2, GPU session 13: prompt prefill (tests + timing), the EnCodec bench workload, ncu evidence (launch list + DRAM bytes of the
# default step at KV 751; full-set capture of the fused step kernel).
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== prefill tests"; $T 300 python -m pytest tests/test_gpu_lm.py -q -m gpu -s -k "prefill or logits_and_tokens" > gpurun_out/r2s13_pytest_prefill.log 2>&1; echo "rc=$?"; grep -E "T0=|passed|failed|Error" gpurun_out/r2s13_pytest_prefill.log | tail -8
echo "== prefill perf"; $T 300 python profiles/perf_prefill.py > gpurun_out/r2s13_perf_prefill.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s13_perf_prefill.log
$T 300 python profiles/perf_prefill.py --batch 1 > gpurun_out/r2s13_perf_prefill_b1.log 2>&1; tail -2 gpurun_out/r2s13_perf_prefill_b1.log
echo "== bench encodec workload (64 x 10 s on one GPU)"; $T 420 python bench.py --workload encodec --batch 64 --steps 2 --warmup 2 > gpurun_out/r2s13_bench_encodec.json 2> gpurun_out/r2s13_bench_encodec.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2s13_bench_encodec.json; tail -3 gpurun_out/r2s13_bench_encodec.err
echo "== ncu: launch list + DRAM bytes of the default step at KV 751"
$T 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 700 --csv --log-file gpurun_out/r2_step_kv751_launches_dram.csv python profiles/perf_lm_step.py --one 750 --reps 2 > gpurun_out/r2s13_ncu_step.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s13_ncu_step.log
echo "== ncu: full set of the fused step kernel at KV 751"
ACB_LM_STEP=fused $T 600 ncu --set full --clock-control none --import-source on -k regex:lm_step_kernel -s 1 -c 1 -o gpurun_out/r2_prof_fused_step python profiles/perf_lm_step.py --one 750 --reps 3 > gpurun_out/r2s13_ncu_fused.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s13_ncu_fused.log
ncu -i gpurun_out/r2_prof_fused_step.ncu-rep --page raw --csv > gpurun_out/r2_prof_fused_step_raw.csv 2>/dev/null; ls -la gpurun_out/r2_prof_fused_step* | head
