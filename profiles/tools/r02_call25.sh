#!/bin/bash
# LayerNorm statistics hand-over between GEMMs (producer row sums -> consumer normalising epilogue): parity, bench A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python -m pytest tests/test_zz_lnfold_gpu.py -q -s -k "handover or unet_with" > gpurun_out/r02_c25_tests_ho.log 2>&1; grep -E "^handover|\.handover|unet vs oracle|passed|failed|rror" gpurun_out/r02_c25_tests_ho.log | cut -c1-220 | head -30
if grep -q "failed\|rror" gpurun_out/r02_c25_tests_ho.log || ! grep -q passed gpurun_out/r02_c25_tests_ho.log; then echo "hand-over failing: VX_LN_FUSE=0 for the rest"; tail -40 gpurun_out/r02_c25_tests_ho.log | cut -c1-200; export VX_LN_FUSE=0; fi
echo "== op tests done at $(( $(date +%s) - S )) s"
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_fullwidth_gpu.py tests/test_pipeline_gpu.py tests/test_zz_gemm_pairs_gpu.py -q -s > gpurun_out/r02_c25_tests_b.log 2>&1; grep -E "rel|err|tap .*e-0[12]" gpurun_out/r02_c25_tests_b.log | grep -v "^tap" | cut -c1-200 | tail -12; tail -3 gpurun_out/r02_c25_tests_b.log | cut -c1-300
echo "== tests done at $(( $(date +%s) - S )) s"
for v in 0 1 0 1; do
[ "$VX_LN_FUSE" = 0 ] && [ $v = 1 ] && continue
VX_LN_FUSE=$v VX_BENCH_NO_CPU=1 VX_BENCH_OPS=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c25_bench_fuse$v.json 2> gpurun_out/r02_c25_bench_fuse$v.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_c25_bench_fuse$v.json").read().strip().splitlines()[-1])
    print("VX_LN_FUSE=$v", d["value"], d["e2e"]["value"], d.get("unet_ms_per_step"), d.get("vae_decode_ms"), d.get("clocks"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r02_c25_bench_fuse$v.err").read()[-1500:])
PY
done
grep "by op" gpurun_out/r02_c25_bench_fuse0.err | head -1; grep "by op" gpurun_out/r02_c25_bench_fuse1.err | head -1
grep -E "gemm_rowsums|gemm_lnparts" gpurun_out/r02_c25_bench_fuse1.err | head -24 | cut -c1-170
echo "== all done at $(( $(date +%s) - S )) s"
