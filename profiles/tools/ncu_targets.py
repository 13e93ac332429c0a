"""Small driver for `ncu --set full` captures of single kernels at the UNet's level-0 shapes (one launch each after a warm-up).
   ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/<name> python profiles/tools/ncu_targets.py <what>"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import ops
what = sys.argv[1]
torch.manual_seed(0)
M, C = 131072, 320
if what == "flash":
    qkv = torch.randn(M, 3 * C, device="cuda").bfloat16()
    for _ in range(3):
        ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], 8, 4096, 4096)
elif what == "groupnorm":
    x = torch.randn(M, C, device="cuda").bfloat16()
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    for _ in range(3):
        ops.groupnorm(x, 32, 4096, g, b, 1e-5, True)
elif what == "layernorm":
    x = torch.randn(M, C, device="cuda").bfloat16()
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    for _ in range(3):
        ops.layernorm(x, g, b)
elif what == "gemm320":
    a = torch.randn(M, C, device="cuda").bfloat16()
    w = (torch.randn(C, C, device="cuda") / 18).bfloat16()
    bias = torch.zeros(C, device="cuda")
    res = torch.randn(M, C, device="cuda").bfloat16()
    for _ in range(3):
        ops.gemm(a, w, bias, residual=res)
elif what == "gemm_qkv":       # level-0 fused QKV projection: N = 960, no residual -- neither MMA-, HBM- nor issue-bound on paper
    a = torch.randn(M, C, device="cuda").bfloat16()
    w = (torch.randn(3 * C, C, device="cuda") / 18).bfloat16()
    out = torch.empty(M, 3 * C, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, None, out=out)
elif what == "gemm_k640":      # level-1 out-projection: M = 32768, K = N = 640, + residual
    a = torch.randn(32768, 640, device="cuda").bfloat16()
    w = (torch.randn(640, 640, device="cuda") / 25).bfloat16()
    bias = torch.zeros(640, device="cuda")
    res = torch.randn(32768, 640, device="cuda").bfloat16()
    for _ in range(3):
        ops.gemm(a, w, bias, residual=res)
elif what == "gemm_ff1":
    a = torch.randn(M, C, device="cuda").bfloat16()
    w = (torch.randn(8 * C, C, device="cuda") / 18).bfloat16()
    wp, bp, _ = ops.pack_geglu(w, torch.zeros(8 * C, device="cuda"))
    for _ in range(3):
        ops.gemm(a, wp, bp, geglu=True)
torch.cuda.synchronize()
