#!/bin/bash
# 1 GPU: BASELINE configs[2] (96 frames, 11 windows) and configs[4] (768x768, 16 frames, 50 steps)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
VX_BENCH_NO_CPU=1 timeout 900 python bench.py --frames 96 --steps 1 --warmup 3 > gpurun_out/r02_c3_96frames.json 2> gpurun_out/r02_c3_96frames.err
VX_BENCH_NO_CPU=1 timeout 900 python bench.py --size 768 --ddim-steps 50 --steps 1 --warmup 3 > gpurun_out/r02_c5_768.json 2> gpurun_out/r02_c5_768.err
cut -c1-300 gpurun_out/r02_c3_96frames.json; tail -2 gpurun_out/r02_c3_96frames.err; cut -c1-300 gpurun_out/r02_c5_768.json; tail -3 gpurun_out/r02_c5_768.err
