#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -s -k "flash" > gpurun_out/r02_c4_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_c4_tests.log
timeout 600 python profiles/tools/fa_sweep.py > gpurun_out/r02_c4_fa_sweep.txt 2>&1
# roofline metrics of one eager UNet forward + VAE decode
timeout 1500 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv --log-file gpurun_out/r02_roofline_raw.csv python profiles/tools/forward_once.py gpurun_out/r02_oplog.json > gpurun_out/r02_c4_ncu.log 2>&1
python profiles/tools/roofline_merge.py gpurun_out/r02_roofline_raw.csv gpurun_out/r02_oplog.json gpurun_out/r02_roofline.csv >> gpurun_out/r02_c4_ncu.log 2>&1
tail -3 gpurun_out/r02_c4_tests.log; cat gpurun_out/r02_c4_fa_sweep.txt; tail -16 gpurun_out/r02_c4_ncu.log
