#!/bin/bash
# why do the K = 320 / 640 GEMMs run at 3x their MMA time?  tile / ring sweep + ncu --set full of the QKV projection
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python profiles/tools/gemm_sweep.py auto bn64 bn128 bn160 bn256 nbuf1 st2 st4 2>&1 | grep -v "^\[vx_gemm\]" | head -14 | tee gpurun_out/r02_c26_gemm_sweep.txt
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/r02_gemm_qkv python profiles/tools/ncu_targets.py gemm_qkv > gpurun_out/r02_c26_ncu1.log 2>&1; tail -2 gpurun_out/r02_c26_ncu1.log
timeout 300 $N -k regex:gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/r02_gemm_k640 python profiles/tools/ncu_targets.py gemm_k640 > gpurun_out/r02_c26_ncu2.log 2>&1; tail -2 gpurun_out/r02_c26_ncu2.log
ls -la gpurun_out/*.ncu-rep | tail -3
