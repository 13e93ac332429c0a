#!/bin/bash
# (a) pair accumulator hand-back without the cluster-scope fence: parity, sweep, bench A/B (+ pairs from K = 320 on top);
# (b) does W multicast / a CTA pair reduce the L2 slice traffic of the level-0 QKV GEMM?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_zz_gemm_pairs_gpu.py tests/test_ops_gpu.py -q -x > gpurun_out/r02_c28_tests_a.log 2>&1; tail -3 gpurun_out/r02_c28_tests_a.log | cut -c1-200
echo "== tests done at $(( $(date +%s) - S )) s"
python - <<'PY' 2>&1 | grep -v "^\[vx_gemm\]" | tee gpurun_out/r02_c28_gemm_sweep.txt
import os, sys, runpy
sys.argv = ["gemm_sweep.py", "auto", "strict", "cg2", "cg2 strict", "minkb5"]
src = open("profiles/tools/gemm_sweep.py").read()
src = src.replace('settings = [', 'settings = [("strict", {"VX_GEMM_STRICT_ARRIVE": "1"}), ("cg2 strict", {"VX_GEMM_CG": "2", "VX_GEMM_STRICT_ARRIVE": "1"}), ("minkb5", {"VX_GEMM_CG_MINKB": "5"}), ', 1)
src = src.replace('keys = [', 'keys = ["VX_GEMM_STRICT_ARRIVE", "VX_GEMM_CG_MINKB", ', 1)
exec(compile(src, "gemm_sweep.py", "exec"))
PY
echo "== sweep done at $(( $(date +%s) - S )) s"
for v in "VX_GEMM_STRICT_ARRIVE=1" "VX_GEMM_STRICT_ARRIVE=0" "VX_GEMM_STRICT_ARRIVE=1" "VX_GEMM_STRICT_ARRIVE=0" "VX_GEMM_CG_MINKB=5"; do
env $v VX_BENCH_NO_CPU=1 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/r02_c28_bench.json 2> gpurun_out/r02_c28_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_c28_bench.json").read().strip().splitlines()[-1])
    print("$v", d["value"], d["e2e"]["value"], d.get("unet_ms_per_step"), d.get("vae_decode_ms"), d.get("clocks"))
except Exception as e:
    print("$v bench failed", e); print(open("gpurun_out/r02_c28_bench.err").read()[-1500:])
PY
done
echo "== benches done at $(( $(date +%s) - S )) s"
M="gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
for v in "VX_GEMM_MC=0" "VX_GEMM_MC=1" "VX_GEMM_CG=2"; do
echo "== $v gemm_qkv"
env $v timeout 200 ncu --clock-control none --metrics $M -k regex:gemm_tcgen05 -s 2 -c 1 python profiles/tools/ncu_targets.py gemm_qkv 2>&1 | grep -E "gpu__time|lts__t|dram__bytes|tensor_cycles" | sed 's/  */ /g'
done 2>&1 | tee gpurun_out/r02_c28_l2_traffic.txt
echo "== all done at $(( $(date +%s) - S )) s"
