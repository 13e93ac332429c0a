"""Each hot op at the UNet's level-0/1/2 shapes, run repeatedly on identical inputs (with an unrelated kernel in between to
disturb the timing) and compared bit for bit with the first run."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import _ffi, ops
torch.manual_seed(0)
dev = 'cuda'


def bf(*s, sc=1.0):
    return (torch.randn(*s, device=dev) * sc).bfloat16()


def check(name, fn, n=12):
    first, bad, worst = None, 0, 0.0
    for i in range(n):
        o = fn()
        if i % 2 == 0:
            torch.empty(1 << 25, device=dev).normal_()
        if first is None:
            first = o.clone()
        elif not torch.equal(o, first):
            bad += 1
            worst = max(worst, (o.float() - first.float()).abs().max().item())
    print(f"{'DIFF' if bad else 'ok  '} {name}: {bad}/{n - 1} runs differ (max |diff| {worst:.3e})", flush=True)


NB = int(sys.argv[1]) if len(sys.argv) > 1 else 8
f = NB // 2
for C, HW, heads in ((320, 4096, 8), (640, 1024, 8), (1280, 256, 8), (1280, 64, 8)):
    M = NB * HW
    x = bf(M, C)
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    check(f"layernorm C={C} M={M}", lambda: ops.layernorm(x, g, b))
    pe = torch.randn(f, C, device=dev)
    check(f"layernorm+pe C={C}", lambda: ops.layernorm(x, g, b, pe=pe, rows_per_frame=HW))
    check(f"groupnorm C={C} HW={HW}", lambda: ops.groupnorm(x, NB, HW, g, b, 1e-6, False))
    check(f"groupnorm+silu C={C} HW={HW}", lambda: ops.groupnorm(x, NB, HW, g, b, 1e-5, True))
    x2 = bf(M, C)
    g2, b2 = torch.randn(2 * C, device=dev), torch.randn(2 * C, device=dev)
    check(f"groupnorm 2-source C={C}+{C}", lambda: ops.groupnorm(x, NB, HW, g2, b2, 1e-5, True, x2=x2))
    w3 = bf(3 * C, C, sc=0.03)
    check(f"gemm qkv K={C} N={3 * C}", lambda: ops.gemm(x, w3))
    w1 = bf(C, C, sc=0.03)
    bias = torch.randn(C, device=dev)
    res = bf(M, C)
    check(f"gemm K={C} N={C} +res", lambda: ops.gemm(x, w1, bias, residual=res))
    check(f"gemm K={C} N={C} +res scale", lambda: ops.gemm(x, w1, bias, residual=res, scale=0.95))
    wg, bg, _ = ops.pack_geglu(bf(8 * C, C, sc=0.03), torch.randn(8 * C, device=dev))
    check(f"gemm geglu K={C} N={8 * C}", lambda: ops.gemm(x, wg, bg, geglu=True))
    x4 = bf(M, 4 * C)
    w4 = bf(C, 4 * C, sc=0.02)
    check(f"gemm K={4 * C} N={C} +res", lambda: ops.gemm(x4, w4, bias, residual=res))
    qkv = bf(M, 3 * C)
    check(f"flash self hd={C // heads} N={HW}", lambda: ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, HW, HW))
    kv = bf(2 * HW, 2 * C)
    q = bf(M, C)
    check(f"flash bank hd={C // heads} N={HW} kv_div={f}", lambda: ops.flash_attention(q[f * HW:], kv[HW:, :C], kv[HW:, C:], heads, HW, HW, kv_div=f))
    check(f"temporal attention C={C} f={f}", lambda: ops.temporal_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], 2, f, HW, heads))
    kv2 = bf(NB * 5, 2 * C)
    check(f"smallkv attention C={C}", lambda: ops.smallkv_attention(q, kv2[:, :C], kv2[:, C:], HW, heads, 5))
    xi = x.view(NB, int(HW ** 0.5), int(HW ** 0.5), C)
    w9 = bf(C, 9 * C, sc=0.02)
    check(f"im2col_s2 + gemm C={C}", lambda: ops.gemm(ops.im2col_s2(x, NB, int(HW ** 0.5), int(HW ** 0.5)), w9) if HW > 64 else x)
    wu = ops.pack_upconv_weight(bf(C, C, 3, 3, sc=0.02))
    check(f"upconv3x3 C={C}", lambda: ops.upconv3x3(xi, wu, bias))
