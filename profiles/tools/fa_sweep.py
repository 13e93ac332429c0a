"""A/B of the flash-attention kernels on the UNet's shapes: v4 / v3 (one query tile per CTA, double-buffered S, 2 CTAs per SM) against
v2 (two tiles per CTA, 16 softmax warps), plus v3's knobs (share of exponentials on the FMA pipe, K/V ring depth).
Every variant is checked against fp32 SDPA on the first shape.   usage: python profiles/tools/fa_sweep.py"""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.getcwd())
from vexpress_b200 import _ffi, ops
torch.manual_seed(0)
KNOBS = ("VX_FA_V2", "VX_FA_POLY", "VX_FA3_STAGES", "VX_FA_NOONES", "VX_FA3_NOLOAD")


def setenv(**kw):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)
    _ffi.lib().vx_flash_reload_env()


def run(B, N, Nk, heads, hd, kv_div, label, check=False):
    C = heads * hd
    q = torch.randn(B * N, C, device='cuda').bfloat16()
    kv = torch.randn((B // kv_div) * Nk, 2 * C, device='cuda').bfloat16()
    k, v = kv[:, :C], kv[:, C:]
    flops = 4.0 * B * heads * N * Nk * hd

    def t_ms(n=5):
        for _ in range(2):
            o = ops.flash_attention(q, k, v, heads, N, Nk, kv_div=kv_div)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            o = ops.flash_attention(q, k, v, heads, N, Nk, kv_div=kv_div)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, o
    ref = None
    if check:
        qf = q.float().view(B, N, heads, hd).transpose(1, 2)
        kf = k.float().reshape(B // kv_div, Nk, heads, hd).transpose(1, 2).repeat_interleave(kv_div, 0)
        vf = v.float().reshape(B // kv_div, Nk, heads, hd).transpose(1, 2).repeat_interleave(kv_div, 0)
        ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * N, C)
    print(label)
    for name, kw in (("v3 (default)", {}), ("v3 4 stages", dict(VX_FA3_STAGES=4)),
                     ("v3, K/V traffic removed (timing experiment, wrong results)", dict(VX_FA3_NOLOAD=1)),
                     ("v3 poly 1/4", dict(VX_FA_POLY=4)), ("v2", dict(VX_FA_V2=1))):
        setenv(**kw)
        try:
            ms, o = t_ms()
        except Exception as e:
            print(f"  {name:62s} failed: {e}")
            continue
        err = "" if (ref is None or "NOLOAD" in str(kw)) else f"  rel-L2 vs fp32 SDPA {((o.float() - ref).norm() / ref.norm()).item():.2e}"
        print(f"  {name:62s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s{err}", flush=True)
    setenv()


run(4, 4096, 4096, 8, 40, 1, "check (B=4 N=4096 hd=40)", check=True)
run(32, 4096, 4096, 8, 40, 1, "level-0 self (B=32 N=4096 hd=40)")
run(16, 4096, 4096, 8, 40, 16, "level-0 bank (B=16 N=4096 hd=40 kv_div=16)")
run(32, 1024, 1024, 8, 80, 1, "level-1 self (B=32 N=1024 hd=80)")
run(16, 1024, 1024, 8, 80, 16, "level-1 bank (B=16 N=1024 hd=80 kv_div=16)")
run(32, 256, 256, 8, 160, 1, "level-2 self (B=32 N=256 hd=160)")
