#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -s -k "flash" > gpurun_out/r02_c3_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c3_tests.log
timeout 600 python profiles/tools/fa_sweep.py > gpurun_out/r02_c3_fa_sweep.txt 2>&1
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_fullwidth_gpu.py -q -s -rfEs > gpurun_out/r02_c3_tests2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c3_tests2.log
VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c3_bench.json 2> gpurun_out/r02_c3_bench.err
tail -3 gpurun_out/r02_c3_tests.log; cat gpurun_out/r02_c3_fa_sweep.txt; tail -3 gpurun_out/r02_c3_tests2.log; cut -c1-300 gpurun_out/r02_c3_bench.json
