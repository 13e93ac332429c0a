"""Sweep of the softmax scheduling knobs (VX_FA_PAIRSYNC: 64-thread pair barrier for the max exchange) and of the two softmax scheduling knobs of flash_attn2_kernel: start offset of query tile 1 (VX_FA_STAGGER, clocks)
and waiting for P.V(j-1) only before the P store (VX_FA_LATEWAIT).   usage: python profiles/tools/fa_stagger.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import ops
torch.manual_seed(0)

def run(B, N, Nk, heads, hd, kv_div, label):
    C = heads * hd
    q = torch.randn(B * N, C, device='cuda').bfloat16()
    kv = torch.randn((B // kv_div) * Nk, 2 * C, device='cuda').bfloat16()
    k, v = kv[:, :C], kv[:, C:]
    def t_ms(n=5):
        for _ in range(2): o = ops.flash_attention(q, k, v, heads, N, Nk, kv_div=kv_div)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): o = ops.flash_attention(q, k, v, heads, N, Nk, kv_div=kv_div)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, o
    for k_ in ("VX_FA_STAGGER", "VX_FA_LATEWAIT", "VX_FA_PAIRSYNC"): os.environ.pop(k_, None)
    base_ms, base = t_ms()
    print(f"{label}: baseline {base_ms:.3f} ms")
    for lw in (0, 1, 2, 3):
        row = f"  latewait={lw & 1} pairsync={lw >> 1}:"
        for st in (0, 1300):
            os.environ["VX_FA_STAGGER"] = str(st); os.environ["VX_FA_LATEWAIT"] = str(lw & 1)
            os.environ["VX_FA_PAIRSYNC"] = str(lw >> 1)
            ms, o = t_ms()
            d = (o.float() - base.float()).abs().max().item()
            row += f"  st{st}={ms:.3f}" + ("" if d == 0 else f"(d={d:.1e})")
        print(row, flush=True)
    for k_ in ("VX_FA_STAGGER", "VX_FA_LATEWAIT", "VX_FA_PAIRSYNC"): os.environ.pop(k_, None)
    # exponential-phase baton between the two query tiles (VX_FA_BATON, template instantiation of its own)
    for lw in (0, 1):
        for poly in (8, 4, 2):             # 1/poly of the exponentials on the FMA pipe (hd <= 64 kernels only)
            os.environ["VX_FA_BATON"] = "1"; os.environ["VX_FA_LATEWAIT"] = str(lw); os.environ["VX_FA_POLY"] = str(poly)
            ms, o = t_ms()
            d = (o.float() - base.float()).abs().max().item()
            print(f"  baton=1 latewait={lw} poly=1/{poly}: {ms:.3f} ms" + ("" if d == 0 else f" (max abs diff vs baseline {d:.1e})"), flush=True)
    for k_ in ("VX_FA_BATON", "VX_FA_LATEWAIT", "VX_FA_POLY"): os.environ.pop(k_, None)

run(32, 4096, 4096, 8, 40, 1, "level-0 self (B=32 N=4096 hd=40)")
run(16, 4096, 4096, 8, 40, 16, "level-0 bank (B=16 N=4096 hd=40 kv_div=16)")
run(32, 1024, 1024, 8, 80, 1, "level-1 self (B=32 N=1024 hd=80)")
