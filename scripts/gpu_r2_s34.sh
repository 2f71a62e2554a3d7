#!/bin/bash
# Round 2, GPU session 34 (2 GPUs): weak- and strong-scaling bench lines + the reference arm under torchrun; world-size-2 GPU test.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== weak"; $T 500 $R --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-ref-gpu > gpurun_out/r2s34_bench_2gpu_weak.json 2> gpurun_out/r2s34_bench_2gpu_weak.err; echo "rc=$?"; cut -c1-500 gpurun_out/r2s34_bench_2gpu_weak.json
echo "== strong"; $T 500 $R --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 3 --no-ref-gpu --scaling strong > gpurun_out/r2s34_bench_2gpu_strong.json 2> gpurun_out/r2s34_bench_2gpu_strong.err; echo "rc=$?"; cut -c1-700 gpurun_out/r2s34_bench_2gpu_strong.json
echo "== encodec strong"; $T 400 $R --master-port 29513 bench.py --gpus 2 --workload encodec --batch 64 --steps 2 --warmup 2 --scaling strong > gpurun_out/r2s34_bench_2gpu_encodec.json 2> gpurun_out/r2s34_bench_2gpu_encodec.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2s34_bench_2gpu_encodec.json
echo "== reference arm under torchrun"; $T 300 $R --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2s34_bench_ref_2gpu.json 2> gpurun_out/r2s34_bench_ref_2gpu.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2s34_bench_ref_2gpu.json
