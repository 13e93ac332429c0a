#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 ./profiles/tools/build/ubench > gpurun_out/r02_c2_ubench.txt 2>&1
timeout 1500 python -m pytest tests/test_pipeline_gpu.py tests/test_fullwidth_gpu.py tests/test_zz_lnfold_gpu.py tests/test_unet_gpu.py -q -s -rfEs > gpurun_out/r02_c2_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c2_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_c2_smoke.txt 2>&1
VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err
tail -4 gpurun_out/r02_c2_tests.log; tail -25 gpurun_out/r02_c2_ubench.txt; cut -c1-400 gpurun_out/r02_c2_bench.json
