#!/bin/bash
# flash v3 ordering fix: determinism sweep, timing sweep, whole GPU suite, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python profiles/tools/flash_determinism.py > gpurun_out/r02_c15_fadet.txt 2>&1; grep -c "^ok" gpurun_out/r02_c15_fadet.txt; grep -v "^ok" gpurun_out/r02_c15_fadet.txt | cut -c1-200
timeout 300 python profiles/tools/fa_sweep.py > gpurun_out/r02_c15_fa_sweep.txt 2>&1; cut -c1-160 gpurun_out/r02_c15_fa_sweep.txt
timeout 300 python profiles/tools/op_determinism.py 8 > gpurun_out/r02_c15_opdet.txt 2>&1; grep -c "^ok" gpurun_out/r02_c15_opdet.txt; grep -v "^ok" gpurun_out/r02_c15_opdet.txt | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_c15_tests.log 2>&1; tail -5 gpurun_out/r02_c15_tests.log
VX_BENCH_NO_CPU=1 VX_BENCH_OPS=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c15_bench.json 2> gpurun_out/r02_c15_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c15_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d.get("unet_ms_per_step"), d.get("vae_decode_ms"), d.get("clocks"))
PY
grep "by op" gpurun_out/r02_c15_bench.err
