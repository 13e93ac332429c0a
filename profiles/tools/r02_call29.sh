#!/bin/bash
# final validation of the round: whole GPU suite, smoke(), driver-form bench, ncu roofline pass + launch list of the final kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_c29_tests.log 2>&1; tail -14 gpurun_out/r02_c29_tests.log | cut -c1-200
echo "== tests done at $(( $(date +%s) - S )) s"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c29_smoke.log 2>&1; grep -E "smoke|Error|rror" gpurun_out/r02_c29_smoke.log | cut -c1-200
echo "== smoke done at $(( $(date +%s) - S )) s"
timeout 900 python bench.py > gpurun_out/r02_c29_bench_default.json 2> gpurun_out/r02_c29_bench_default.err; tail -1 gpurun_out/r02_c29_bench_default.json | cut -c1-700
echo "== default bench done at $(( $(date +%s) - S )) s"
VX_BENCH_NO_CPU=1 VX_BENCH_OPS=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c29_bench_ops.json 2> gpurun_out/r02_c29_bench_ops.err; grep "by op" gpurun_out/r02_c29_bench_ops.err | cut -c1-400
echo "== op table done at $(( $(date +%s) - S )) s"
timeout 1200 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv --log-file gpurun_out/r02_roofline_raw.csv python profiles/tools/forward_once.py gpurun_out/r02_oplog.json > gpurun_out/r02_c29_ncu.log 2>&1
python profiles/tools/roofline_merge.py gpurun_out/r02_roofline_raw.csv gpurun_out/r02_oplog.json gpurun_out/r02_roofline.csv >> gpurun_out/r02_c29_ncu.log 2>&1
grep -E "launches logged|WARNING|rows ->" gpurun_out/r02_c29_ncu.log | cut -c1-250
python profiles/tools/summarize_launches.py gpurun_out/r02_roofline_raw.csv > gpurun_out/r02_launch_summary.csv 2>/dev/null; head -12 gpurun_out/r02_launch_summary.csv | cut -c1-200
gzip -f gpurun_out/r02_roofline_raw.csv
echo "== all done at $(( $(date +%s) - S )) s"
