#!/bin/bash
# round-2 first GPU session: everything that had never run + the prepared flash/LN-fold experiments
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_c1_gpu.txt 2>&1
nproc >> gpurun_out/r02_c1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s -rfEs > gpurun_out/r02_c1_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c1_tests.log
timeout 120 ./profiles/tools/build/ubench > gpurun_out/r02_c1_ubench.txt 2>&1
timeout 400 python profiles/tools/fa_stagger.py > gpurun_out/r02_c1_fa_stagger.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_c1_smoke.txt 2>&1
echo "smoke exit $?" >> gpurun_out/r02_c1_smoke.txt
VX_BENCH_NO_CPU=1 VX_BENCH_OPS=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c1_bench_default.json 2> gpurun_out/r02_c1_bench_default.err
VX_LN_FOLD=1 VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c1_bench_lnfold.json 2> gpurun_out/r02_c1_bench_lnfold.err
VX_FA_BATON=1 VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c1_bench_baton.json 2> gpurun_out/r02_c1_bench_baton.err
tail -5 gpurun_out/r02_c1_tests.log
cat gpurun_out/r02_c1_bench_default.json | cut -c1-600
