"""Bring-up probe: flash v3 with the MMA warp delayed before every P.V (VX_FA3_DBG bit 5) at several head dims / lengths,
alone and combined with the other ordering switches.  Prints rel-L2 error vs fp32 SDPA."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import _ffi, ops
torch.manual_seed(0)
heads = 4
for hd, N, B in ((40, 128, 2), (40, 192, 2), (40, 256, 2), (40, 1024, 4), (64, 128, 2), (64, 1024, 4), (80, 128, 2), (80, 192, 2), (80, 256, 2), (80, 1024, 4)):
    C = heads * hd
    qkv = torch.randn(B * N, 3 * C, device='cuda').bfloat16()
    ref = torch.nn.functional.scaled_dot_product_attention(*[qkv[:, i * C:(i + 1) * C].float().view(B, N, heads, hd).transpose(1, 2) for i in range(3)])
    ref = ref.transpose(1, 2).reshape(B * N, C)
    row = f"hd={hd:3d} N={N:4d}:"
    for dbg in (0, 32, 33, 36, 40, 34, 48):
        os.environ["VX_FA3_DBG"] = str(dbg)
        _ffi.lib().vx_flash_reload_env()
        o = ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, N, N).float()
        row += f"  dbg{dbg}={((o - ref).norm() / ref.norm()).item():.2e}"
    print(row, flush=True)
