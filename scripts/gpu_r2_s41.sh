#!/bin/bash
# Round 2, GPU session 41: end-of-round tree -- every GPU test, smoke, bench (without the 2-minute reference-CUDA pass, measured in
# session 33), EnCodec bench workload.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== all GPU tests"; $T 1500 python -m pytest tests -q -m gpu > gpurun_out/r2s41_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s41_pytest_gpu.log
echo "== smoke"; $T 300 python __graft_entry__.py smoke > gpurun_out/r2s41_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s41_smoke.log
echo "== bench encodec workload"; $T 420 python bench.py --workload encodec --batch 64 --steps 2 --warmup 2 > gpurun_out/r2s41_bench_encodec.json 2> gpurun_out/r2s41_bench_encodec.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2s41_bench_encodec.json
echo "== bench b200"; $T 700 python bench.py --steps 1 --warmup 3 --no-ref-gpu > gpurun_out/r2s41_bench.json 2> gpurun_out/r2s41_bench.err; echo "rc=$?"; cut -c1-500 gpurun_out/r2s41_bench.json
