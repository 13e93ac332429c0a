#!/bin/bash
# LayerNorm + positional encoding through layernorm5_kernel: parity, A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -s -k "layernorm" > gpurun_out/r02_c31_tests_ln.log 2>&1; grep -E "layernorm rows|passed|failed|rror" gpurun_out/r02_c31_tests_ln.log | cut -c1-200
if grep -q "failed\|rror" gpurun_out/r02_c31_tests_ln.log; then echo "LN+PE failing"; export VX_LN_PE5=0; fi
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fullwidth_gpu.py tests/test_pipeline_gpu.py tests/test_zz_pdl_gpu.py -q > gpurun_out/r02_c31_tests_b.log 2>&1; tail -3 gpurun_out/r02_c31_tests_b.log | cut -c1-300
for v in 0 1 0 1; do
VX_LN_PE5=$v VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c31_bench.json 2> gpurun_out/r02_c31_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_c31_bench.json").read().strip().splitlines()[-1])
    print("VX_LN_PE5=$v", d["value"], d["e2e"]["value"], d.get("unet_ms_per_step"), d.get("vae_decode_ms"), d.get("clocks"), d["roofline"]["layernorm"]["seconds_per_forward"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_c31_bench.err").read()[-1500:])
PY
done
