#!/bin/bash
# Round 2, GPU session 35: the tree with the ex2.approx ELU as default -- every GPU test (incl. the new attention bit-identity test),
# smoke, EnCodec bench workload.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== all GPU tests"; $T 1500 python -m pytest tests -q -m gpu > gpurun_out/r2s35_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s35_pytest_gpu.log
echo "== smoke"; $T 300 python __graft_entry__.py smoke > gpurun_out/r2s35_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s35_smoke.log
echo "== bench encodec workload"; $T 420 python bench.py --workload encodec --batch 64 --steps 2 --warmup 2 > gpurun_out/r2s35_bench_encodec.json 2> gpurun_out/r2s35_bench_encodec.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2s35_bench_encodec.json
echo "== encodec perf"; $T 300 python profiles/perf_encodec.py > gpurun_out/r2s35_perf_encodec.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2s35_perf_encodec.log
