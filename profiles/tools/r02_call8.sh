#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python profiles/tools/fa_sweep.py > gpurun_out/r02_c8_fa_sweep.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -rfEs > gpurun_out/r02_c8_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c8_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_c8_smoke.txt 2>&1
VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_c8_bench.json 2> gpurun_out/r02_c8_bench.err
head -14 gpurun_out/r02_c8_fa_sweep.txt; tail -6 gpurun_out/r02_c8_tests.log; tail -4 gpurun_out/r02_c8_smoke.txt; cut -c1-250 gpurun_out/r02_c8_bench.json
