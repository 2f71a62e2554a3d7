#!/bin/bash
# Round 2, GPU session 20: per-phase step with the software-pipelined GEMM k-loop (v10 cluster experiment removed), full GPU test
# files, EnCodec bench workload with the fused residual blocks.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== step perf"; $T 400 python profiles/perf_lm_step.py > gpurun_out/r2s20_perf_step.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s20_perf_step.log
echo "== LM + fullsize tests"; $T 1200 python -m pytest tests/test_gpu_lm.py tests/test_gpu_fullsize.py tests/test_gpu_dist.py -q -m gpu > gpurun_out/r2s20_pytest_lm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s20_pytest_lm.log
echo "== bench encodec workload"; $T 420 python bench.py --workload encodec --batch 64 --steps 2 --warmup 2 > gpurun_out/r2s20_bench_encodec.json 2> gpurun_out/r2s20_bench_encodec.err; echo "rc=$?"; cut -c1-600 gpurun_out/r2s20_bench_encodec.json
