#!/bin/bash
# Round 2, GPU session 12: EnCodec with the fp32_tc encoder default and 128B-swizzled conv1d_t5 tiles; full GPU test-suite
# (incl. full-size goldens, fused opt-in, RoPE, state API); bench with the reference arms.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
echo "== encodec tests"; $T 600 python -m pytest tests/test_gpu_encodec.py -q -m gpu > gpurun_out/r2s12_pytest_encodec.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2s12_pytest_encodec.log
echo "== encodec perf (defaults)"; $T 300 python profiles/perf_encodec.py --enc fp32_tc > gpurun_out/r2s12_perf_encodec.log 2>&1; echo "rc=$?"; tail -34 gpurun_out/r2s12_perf_encodec.log
echo "== LM + fullsize + dist tests"; $T 900 python -m pytest tests/test_gpu_lm.py tests/test_gpu_fullsize.py -q -m gpu > gpurun_out/r2s12_pytest_lm.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2s12_pytest_lm.log
echo "== smoke"; $T 300 python __graft_entry__.py smoke > gpurun_out/r2s12_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s12_smoke.log
echo "== bench b200"; $T 600 python bench.py --steps 1 --warmup 3 > gpurun_out/r2s12_bench.json 2> gpurun_out/r2s12_bench.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2s12_bench.json; grep reference_arm gpurun_out/r2s12_bench.err | tail -12
echo "== bench reference arm"; $T 420 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s12_bench_ref.json 2> gpurun_out/r2s12_bench_ref.err; echo "rc=$?"; cut -c1-1200 gpurun_out/r2s12_bench_ref.json; grep reference_arm gpurun_out/r2s12_bench_ref.err | tail -12
