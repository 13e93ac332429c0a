#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py -q -s -k "groupnorm or upconv or conv" > gpurun_out/r02_c5_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c5_tests.log
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py tests/test_fullwidth_gpu.py tests/test_zz_refnet_gpu.py -q -s -rfEs > gpurun_out/r02_c5_tests2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c5_tests2.log
VX_BENCH_NO_CPU=1 VX_BENCH_OPS=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c5_bench.json 2> gpurun_out/r02_c5_bench.err
VX_GN_FUSED=0 VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c5_bench_gn2k.json 2> /dev/null
grep -E "passed|failed|exit" gpurun_out/r02_c5_tests.log gpurun_out/r02_c5_tests2.log; cut -c1-250 gpurun_out/r02_c5_bench.json; grep "by op" gpurun_out/r02_c5_bench.err
