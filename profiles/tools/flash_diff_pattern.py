"""Where do two runs of the same flash attention differ?  Histograms of the differing output elements by head-dim column,
row inside the 128-row query tile, query tile, head and batch; and against the fp32 reference which of the two is wrong."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import _ffi, ops
torch.manual_seed(0)
dev = 'cuda'
heads = 8
hd, N, B = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "80,1024,32").split(","))
for dbg in (sys.argv[2:] or ["0", "2", "1"]):
    os.environ["VX_FA3_DBG"] = dbg
    _ffi.lib().vx_flash_reload_env()
    C = heads * hd
    qkv = torch.randn(B * N, 3 * C, device=dev).bfloat16()
    ref = torch.nn.functional.scaled_dot_product_attention(*[qkv[:, i * C:(i + 1) * C].float().view(B, N, heads, hd).transpose(1, 2) for i in range(3)])
    ref = ref.transpose(1, 2).reshape(B * N, C)
    outs = []
    for i in range(4):
        outs.append(ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, N, N).float())
        torch.empty(1 << 25, device=dev).normal_()
    # element-wise: an element is "bad" in run i if it is further from the reference than 0.02
    print(f"==== hd={hd} N={N} B={B} VX_FA3_DBG={dbg}")
    for i, o in enumerate(outs):
        bad = (o - ref).abs() > 0.02
        nb = int(bad.sum())
        if nb == 0:
            print(f"run {i}: clean (max err {(o - ref).abs().max().item():.3e})")
            continue
        idx = bad.nonzero()
        rows, cols = idx[:, 0], idx[:, 1]
        b_, n_ = rows // N, rows % N
        qt, r = n_ // 128, n_ % 128
        h_, d_ = cols // hd, cols % hd
        def hist(x, n):
            return torch.bincount(x, minlength=n).tolist()
        print(f"run {i}: {nb} bad elements, max err {(o - ref).abs().max().item():.3e}")
        print(f"   by head-dim column /8: {hist(d_ // 8, hd // 8)}")
        print(f"   by row in tile /8  : {hist(r // 8, 16)}")
        if os.environ.get("VX_PATTERN_FULL"):
            print(f"   by query tile      : {hist(qt, N // 128)}")
            print(f"   by head            : {hist(h_, heads)}")
            print(f"   by batch           : {hist(b_, B)}")
        # per bad (batch, head, qtile) CTA: how many rows are bad, and are whole rows bad?
        cta = (b_ * heads + h_) * (N // 128) + qt
        u, cnt = torch.unique(cta, return_counts=True)
        print(f"   bad CTAs: {len(u)} of {B * heads * N // 128}; bad elements per bad CTA: min {int(cnt.min())} median {int(cnt.median())} max {int(cnt.max())}")
        rowkey = cta * 128 + r
        ur, rc = torch.unique(rowkey, return_counts=True)
        print(f"   bad rows: {len(ur)}; bad columns per bad row: min {int(rc.min())} median {int(rc.median())} max {int(rc.max())} (hd = {hd})")
