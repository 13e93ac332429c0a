"""HBM-bound helper kernels in isolation: temporal attention, LayerNorm, GroupNorm at the three UNet levels.
Prints time and algorithmic GB/s (bytes that must move once) for the current kernels and, where an env switch
exists, the previous ones.   usage: python profiles/tools/small_ops_bench.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import ops
torch.manual_seed(0)
dev = 'cuda'
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

def t_us(fn, n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()                      # evict L2 so every run streams from HBM like it does inside the UNet
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3

def row(name, nbytes, fn, env=None):
    cols = []
    for setting in ([None] + ([env] if env else [])):
        if setting: os.environ[setting] = "1"
        us = t_us(fn)
        if setting: os.environ.pop(setting)
        cols.append(f"{us:8.1f} us {nbytes / us * 1e-3:7.0f} GB/s" + (f" ({setting})" if setting else ""))
    print(f"{name:44s}" + "   |   ".join(cols), flush=True)

b, f = 2, 16
for HW, C, heads in ((4096, 320, 8), (1024, 640, 8), (256, 1280, 8)):
    rows = b * f * HW
    qkv = torch.randn(rows, 3 * C, device=dev).bfloat16()
    out = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    row(f"temporal_attention rows={rows} C={C}", rows * C * 2 * 4,
        lambda: ops.temporal_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], b, f, HW, heads, out=out), "VX_TEMPORAL_V1")
    x = torch.randn(rows, C, device=dev).bfloat16()
    g, be = torch.randn(C, device=dev), torch.randn(C, device=dev)
    pe = torch.randn(24, C, device=dev)
    row(f"layernorm rows={rows} C={C}", rows * C * 2 * 2, lambda: ops.layernorm(x, g, be, out=out), "VX_LN_V1")
    row(f"layernorm+pe rows={rows} C={C}", rows * C * 2 * 2,
        lambda: ops.layernorm(x, g, be, pe=pe, rows_per_frame=HW, out=out), "VX_LN_V1")
    row(f"groupnorm+silu NB={b * f} HW={HW} C={C}", rows * C * 2 * 3,
        lambda: ops.groupnorm(x, b * f, HW, g, be, 1e-5, True, out=out))
    x2 = torch.randn(rows, C, device=dev).bfloat16()
    g2, be2 = torch.randn(2 * C, device=dev), torch.randn(2 * C, device=dev)
    out2 = torch.empty(rows, 2 * C, device=dev, dtype=torch.bfloat16)
    row(f"groupnorm+silu concat C={C}+{C}", rows * C * 2 * 6,
        lambda: ops.groupnorm(x, b * f, HW, g2, be2, 1e-5, True, x2=x2, out=out2))
    q = torch.randn(rows, C, device=dev).bfloat16()
    kv = torch.randn(b * f * 5, 2 * C, device=dev).bfloat16()
    row(f"smallkv_attention rows={rows} C={C}", rows * C * 2 * 2,
        lambda: ops.smallkv_attention(q, kv[:, :C], kv[:, C:], HW, heads, 5, out=out))

# conv_in (4 -> 320 channels, 64x64 latents, 32 frames) with the kps-feature addend, as the UNet calls it
xin = torch.randn(32, 4, 64, 64, device=dev).bfloat16()
w = torch.randn(36, 320, device=dev)
bias = torch.randn(320, device=dev)
add = torch.randn(16 * 4096, 320, device=dev).bfloat16()
af = (torch.arange(32, device=dev, dtype=torch.int32) % 16).contiguous()
outc = torch.empty(32 * 4096, 320, device=dev, dtype=torch.bfloat16)
row("conv_in NB=32 64x64 4->320 (+kps addend)", (32 * 4096 * 320 * 2) * 2,
    lambda: ops.conv_in(xin, w, bias, 320, addend=add, add_frame=af, out=outc))
