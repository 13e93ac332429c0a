#!/bin/bash
# 8 GPUs: BASELINE configs[3] (384 frames, 47 windows sharded 6/6/6/6/6/6/6/5) and the weak-scaling line at N = 8
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_mg8_gpus.txt
VX_BENCH_NO_CPU=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --frames 384 --steps 1 --warmup 3 > gpurun_out/r02_c4_384frames_8gpu.json 2> gpurun_out/r02_c4_384frames_8gpu.err
VX_BENCH_NO_CPU=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_mg8_bench.json 2> gpurun_out/r02_mg8_bench.err
cut -c1-500 gpurun_out/r02_c4_384frames_8gpu.json; tail -2 gpurun_out/r02_c4_384frames_8gpu.err; cut -c1-400 gpurun_out/r02_mg8_bench.json
