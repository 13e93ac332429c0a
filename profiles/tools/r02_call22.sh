#!/bin/bash
# stride-2 conv through TMA traversal strides (parity), late-trigger PDL A/B, ncu roofline pass of the final kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -s -k "stride2" > gpurun_out/r02_c22_tests_s2.log 2>&1; grep -E "conv s2|passed|failed|rror" gpurun_out/r02_c22_tests_s2.log | cut -c1-220 | head -30
if grep -q "failed\|rror" gpurun_out/r02_c22_tests_s2.log; then echo "stride-2 conv failing: VX_CONV_S2=0 for the rest"; export VX_CONV_S2=0; fi
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fullwidth_gpu.py tests/test_pipeline_gpu.py tests/test_zz_refnet_gpu.py tests/test_zz_pdl_gpu.py -q > gpurun_out/r02_c22_tests_b.log 2>&1; tail -4 gpurun_out/r02_c22_tests_b.log | cut -c1-300
echo "== tests done at $(( $(date +%s) - S )) s"
for v in 0 1 0 1; do
VX_PDL=$v VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c22_bench_pdl$v.json 2> gpurun_out/r02_c22_bench_pdl$v.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_c22_bench_pdl$v.json").read().strip().splitlines()[-1])
    print("VX_PDL=$v", d["value"], d["e2e"]["value"], d.get("unet_ms_per_step"), d.get("vae_decode_ms"), d.get("clocks"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_c22_bench_pdl$v.err").read()[-1500:])
PY
done
echo "== benches done at $(( $(date +%s) - S )) s"
timeout 1200 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv --log-file gpurun_out/r02_roofline_raw.csv python profiles/tools/forward_once.py gpurun_out/r02_oplog.json > gpurun_out/r02_c22_ncu.log 2>&1
python profiles/tools/roofline_merge.py gpurun_out/r02_roofline_raw.csv gpurun_out/r02_oplog.json gpurun_out/r02_roofline.csv >> gpurun_out/r02_c22_ncu.log 2>&1
tail -20 gpurun_out/r02_c22_ncu.log | cut -c1-200
gzip -f gpurun_out/r02_roofline_raw.csv
echo "== all done at $(( $(date +%s) - S )) s"
