#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/v9_bench_2gpu.json 2> gpurun_out/v9_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-700 gpurun_out/v9_bench_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/v9_bench_ref_2gpu.json 2> gpurun_out/v9_bench_ref_2gpu.err; echo "ref2 rc=$?"; cut -c1-500 gpurun_out/v9_bench_ref_2gpu.json
