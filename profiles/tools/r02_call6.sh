#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:flash_attn3 -s 2 -c 1 -f -o gpurun_out/r02_flash3 python profiles/tools/ncu_targets.py flash > gpurun_out/r02_c6_ncu1.log 2>&1
timeout 300 $N -k regex:gn_fused -s 2 -c 1 -f -o gpurun_out/r02_gn_fused python profiles/tools/ncu_targets.py groupnorm > gpurun_out/r02_c6_ncu2.log 2>&1
timeout 300 $N -k regex:layernorm5 -s 2 -c 1 -f -o gpurun_out/r02_ln5 python profiles/tools/ncu_targets.py layernorm > gpurun_out/r02_c6_ncu3.log 2>&1
timeout 300 $N -k regex:gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/r02_gemm320 python profiles/tools/ncu_targets.py gemm320 > gpurun_out/r02_c6_ncu4.log 2>&1
timeout 300 $N -k regex:gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/r02_gemm_ff1 python profiles/tools/ncu_targets.py gemm_ff1 > gpurun_out/r02_c6_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
