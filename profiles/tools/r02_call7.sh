#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -s -k "groupnorm or layernorm or flash" > gpurun_out/r02_c7_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c7_tests.log
timeout 600 python profiles/tools/fa_sweep.py > gpurun_out/r02_fa_sweep_v4.txt 2>&1
timeout 300 python profiles/tools/small_ops_bench.py > gpurun_out/r02_c7_small_ops.txt 2>&1
VX_LN_BLOCKS=2 timeout 300 python profiles/tools/small_ops_bench.py > gpurun_out/r02_c7_small_ops_ln2.txt 2>&1
VX_BENCH_NO_CPU=1 VX_BENCH_OPS=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c7_bench.json 2> gpurun_out/r02_c7_bench.err
bash profiles/tools/r02_configs_1gpu.sh > gpurun_out/r02_c7_configs.log 2>&1
grep -E "passed|failed|exit" gpurun_out/r02_c7_tests.log; head -18 gpurun_out/r02_fa_sweep_v4.txt; cat gpurun_out/r02_c7_small_ops.txt | tail -12; cut -c1-250 gpurun_out/r02_c7_bench.json; grep "by op" gpurun_out/r02_c7_bench.err; tail -8 gpurun_out/r02_c7_configs.log | cut -c1-300
