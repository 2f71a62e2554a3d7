#!/bin/bash
# Round 2, GPU session 33: full validation of the round's final tree -- every GPU test, smoke, bench (with the reference's own CUDA path
# alongside), the reference arm, the EnCodec bench workload, per-layer EnCodec timing.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== all GPU tests"; $T 1500 python -m pytest tests -q -m gpu > gpurun_out/r2s33_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s33_pytest_gpu.log
echo "== smoke"; $T 300 python __graft_entry__.py smoke > gpurun_out/r2s33_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s33_smoke.log
echo "== encodec perf"; $T 300 python profiles/perf_encodec.py > gpurun_out/r2s33_perf_encodec.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s33_perf_encodec.log
echo "== fast-ELU variant: goldens + perf"; ACB_LIB=$PWD/audiocraft_b200/libaudiocraft_b200_fastelu.so $T 600 python -m pytest tests/test_gpu_encodec.py tests/test_gpu_fullsize.py -q -m gpu -k "golden or properties or resblock or conv1d" > gpurun_out/r2s33_pytest_fastelu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s33_pytest_fastelu.log
ACB_LIB=$PWD/audiocraft_b200/libaudiocraft_b200_fastelu.so $T 300 python profiles/perf_encodec.py > gpurun_out/r2s33_perf_encodec_fastelu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2s33_perf_encodec_fastelu.log
echo "== bench encodec workload"; $T 420 python bench.py --workload encodec --batch 64 --steps 2 --warmup 2 > gpurun_out/r2s33_bench_encodec.json 2> gpurun_out/r2s33_bench_encodec.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2s33_bench_encodec.json
echo "== bench b200"; $T 700 python bench.py --steps 1 --warmup 3 > gpurun_out/r2s33_bench.json 2> gpurun_out/r2s33_bench.err; echo "rc=$?"; cut -c1-700 gpurun_out/r2s33_bench.json
echo "== bench reference arm"; $T 420 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s33_bench_ref.json 2> gpurun_out/r2s33_bench_ref.err; echo "rc=$?"; cut -c1-500 gpurun_out/r2s33_bench_ref.json
