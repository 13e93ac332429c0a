#!/bin/bash
# cluster-resident GroupNorm for the 8x8 / 16x16 levels: parity, timing, A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -s -k "groupnorm" > gpurun_out/r02_c32_tests_gn.log 2>&1; grep -E "groupnorm cluster|passed|failed|rror" gpurun_out/r02_c32_tests_gn.log | cut -c1-220
if grep -q "failed\|rror" gpurun_out/r02_c32_tests_gn.log || ! grep -q passed gpurun_out/r02_c32_tests_gn.log; then echo "cluster GroupNorm failing: stop"; tail -30 gpurun_out/r02_c32_tests_gn.log | cut -c1-200; exit 0; fi
python - <<'PY' 2>&1 | tee gpurun_out/r02_c32_gn_timing.txt
import torch, sys, os
sys.path.insert(0, os.getcwd())
from vexpress_b200 import ops
def t_us(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(f"{'GroupNorm+SiLU shape':36s} {'rendezvous':>12s} {'cluster':>12s}")
for NB, HW, C1, C2 in [(32, 64, 1280, 0), (32, 64, 1280, 1280), (32, 256, 1280, 0), (32, 256, 1280, 640), (32, 256, 1280, 1280), (32, 1024, 640, 0)]:
    x1 = torch.randn(NB * HW, C1, device="cuda").bfloat16()
    x2 = torch.randn(NB * HW, C2, device="cuda").bfloat16() if C2 else None
    g, b = torch.ones(C1 + C2, device="cuda"), torch.zeros(C1 + C2, device="cuda")
    out = torch.empty(NB * HW, C1 + C2, device="cuda", dtype=torch.bfloat16)
    r = []
    for on in (False, True):
        ops._GN_CLUSTER = on
        r.append(t_us(lambda: ops.groupnorm(x1, NB, HW, g, b, 1e-5, True, x2=x2, out=out)))
    print(f"NB={NB} HW={HW} C={C1}+{C2}".ljust(36) + f"{r[0]:12.1f} {r[1]:12.1f}")
PY
VX_GN_CLUSTER=1 timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_fullwidth_gpu.py::test_unet_fullwidth_c1_all_taps tests/test_pipeline_gpu.py tests/test_zz_refnet_gpu.py -q > gpurun_out/r02_c32_tests_b.log 2>&1; tail -3 gpurun_out/r02_c32_tests_b.log | cut -c1-300
for v in 0 1 0 1; do
VX_GN_CLUSTER=$v VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c32_bench.json 2> gpurun_out/r02_c32_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_c32_bench.json").read().strip().splitlines()[-1])
    print("VX_GN_CLUSTER=$v", d["value"], d["e2e"]["value"], d.get("unet_ms_per_step"), d.get("vae_decode_ms"), d.get("clocks"), d["roofline"]["groupnorm"]["seconds_per_forward"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_c32_bench.err").read()[-1500:])
PY
done
