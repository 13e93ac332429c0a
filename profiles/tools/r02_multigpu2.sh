#!/bin/bash
# 2 GPUs: the sharded path equals the single-GPU path (NCCL), then the weak-scaling bench line at N = 2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_mg2_gpus.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -s -rfEs > gpurun_out/r02_mg2_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_mg2_tests.log
VX_BENCH_NO_CPU=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_mg2_bench.json 2> gpurun_out/r02_mg2_bench.err
tail -4 gpurun_out/r02_mg2_tests.log; cut -c1-400 gpurun_out/r02_mg2_bench.json; tail -3 gpurun_out/r02_mg2_bench.err
