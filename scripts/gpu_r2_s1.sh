#!/bin/bash
# Round 2, GPU session 1: grid-barrier floor, first hardware contact of conv1d_t6, full-size parity (golden + live
# reference CUDA path), the reference arms of bench.py.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2s1_gpu.txt 2>&1
echo "== grid barrier"; timeout 120 python profiles/perf_grid_barrier.py > gpurun_out/r2s1_gridbar.log 2>&1; echo "rc=$?"; cat gpurun_out/r2s1_gridbar.log
echo "== t6 smallest"; ACB_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_encodec.py -x -q -s -m gpu \
  -k "experimental_conv1d_t6 and 8-64-7-1-1-False-3" > gpurun_out/r2s1_t6_first.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2s1_t6_first.log
echo "== t6 unit"; ACB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_encodec.py -q -s -m gpu \
  -k "experimental_conv1d_t6" > gpurun_out/r2s1_t6_unit.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2s1_t6_unit.log
echo "== t6 model"; ACB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_encodec.py -q -s -m gpu \
  -k "experimental_flush_encoder" > gpurun_out/r2s1_t6_model.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2s1_t6_model.log
echo "== fullsize parity"; timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -m gpu > gpurun_out/r2s1_fullsize.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2s1_fullsize.log
echo "== bench (b200 arm, with reference_gpu + cpu_baseline)"; timeout 900 python bench.py --steps 1 --warmup 3 > gpurun_out/r2s1_bench.json 2> gpurun_out/r2s1_bench.err; echo "rc=$?"; cat gpurun_out/r2s1_bench.json; tail -5 gpurun_out/r2s1_bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s1_bench_ref.json 2> gpurun_out/r2s1_bench_ref.err; echo "rc=$?"; cat gpurun_out/r2s1_bench_ref.json; tail -3 gpurun_out/r2s1_bench_ref.err
echo "== t6 per-layer timing"; timeout 300 python profiles/perf_encodec.py --enc tf32x3_flush > gpurun_out/r2s1_t6_perf.log 2>&1; tail -40 gpurun_out/r2s1_t6_perf.log
