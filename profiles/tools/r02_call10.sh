#!/bin/bash
# conv row-reuse staging: parity tests, per-shape sweep, bench A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_zz_gemm_pairs_gpu.py tests/test_ops_gpu.py -q -x > gpurun_out/r02_c10_tests_a.log 2>&1; tail -3 gpurun_out/r02_c10_tests_a.log
timeout 600 python profiles/tools/conv_rr_sweep.py > gpurun_out/r02_c10_conv_rr_sweep.txt 2>&1; cat gpurun_out/r02_c10_conv_rr_sweep.txt
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fullwidth_gpu.py tests/test_pipeline_gpu.py -q -x > gpurun_out/r02_c10_tests_b.log 2>&1; tail -3 gpurun_out/r02_c10_tests_b.log
VX_BENCH_NO_CPU=1 VX_BENCH_OPS=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c10_bench_rr1.json 2> gpurun_out/r02_c10_bench_rr1.err
VX_CONV_RR=0 VX_BENCH_NO_CPU=1 VX_BENCH_OPS=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c10_bench_rr0.json 2> gpurun_out/r02_c10_bench_rr0.err
python - <<'PY'
import json
for n in ("rr1", "rr0"):
    try:
        d = json.loads(open(f"gpurun_out/r02_c10_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["e2e"]["value"], d.get("unet_ms_per_step"), d.get("vae_decode_ms"), d.get("clocks"))
        print("   ", json.dumps(d.get("roofline"))[:900])
    except Exception as e:
        print(n, "failed", e)
PY
grep -E 'conv3x3|upconv' gpurun_out/r02_c10_bench_rr1.err | head -40; echo; grep -E 'conv3x3|upconv' gpurun_out/r02_c10_bench_rr0.err | head -40
