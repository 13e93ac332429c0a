#!/bin/bash
# W multicast between 1-CTA tiles (VX_GEMM_MC=1): parity (bit-identical to the plain launch), per-shape sweep, bench A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_zz_gemm_pairs_gpu.py -q -s -k "multicast" > gpurun_out/r02_c24_tests_mc.log 2>&1; grep -E "^mc |\.mc |passed|failed|rror|Timeout" gpurun_out/r02_c24_tests_mc.log | cut -c1-200 | head -30
if grep -q "failed\|rror" gpurun_out/r02_c24_tests_mc.log || ! grep -q passed gpurun_out/r02_c24_tests_mc.log; then echo "multicast failing: stop here"; tail -30 gpurun_out/r02_c24_tests_mc.log | cut -c1-200; exit 0; fi
echo "== tests done at $(( $(date +%s) - S )) s"
timeout 300 python profiles/tools/gemm_sweep.py cg1 mc cg2 2>&1 | grep -v "^\[vx_gemm\]" | head -14 | tee gpurun_out/r02_c24_gemm_sweep.txt
echo "== sweep done at $(( $(date +%s) - S )) s"
VX_GEMM_MC=1 timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_fullwidth_gpu.py::test_unet_fullwidth_c1_all_taps -q > gpurun_out/r02_c24_tests_b.log 2>&1; tail -3 gpurun_out/r02_c24_tests_b.log | cut -c1-300
for v in "VX_GEMM_MC=0" "VX_GEMM_MC=1" "VX_GEMM_CG_MINKB=10" "VX_GEMM_MC=1 VX_GEMM_CG_MINKB=10"; do
env $v VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c24_bench.json 2> gpurun_out/r02_c24_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_c24_bench.json").read().strip().splitlines()[-1])
    print("$v", d["value"], d["e2e"]["value"], d.get("unet_ms_per_step"), d.get("vae_decode_ms"), d.get("clocks"))
except Exception as e:
    print("$v bench failed", e); print(open("gpurun_out/r02_c24_bench.err").read()[-1500:])
PY
done
echo "== all done at $(( $(date +%s) - S )) s"
