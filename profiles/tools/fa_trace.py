"""Bring-up tool: clock64 timeline of one CTA of flash_attn2_kernel (softmax warp 0 and MMA warp of query tile 0)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import ops
torch.manual_seed(0)
B, N, heads, hd = 32, 4096, 8, 40
C = heads * hd
qkv = torch.randn(B * N, 3 * C, device="cuda").bfloat16()
tr = torch.zeros(32 * 16, device="cuda", dtype=torch.int64)
ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, N, N)
os.environ["VX_FA_TRACE"] = str(tr.data_ptr())
ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, N, N)
torch.cuda.synchronize()
t = tr.cpu().view(32, 16)
t0 = int(t[0, 0])
names = {0: "sm:s_full", 1: "sm:ld_done", 2: "sm:xchg", 3: "sm:pv_wait", 6: "sm:exps", 7: "sm:latewait", 4: "sm:st_done", 5: "sm:p_ready",
         8: "mma:kv_full", 9: "mma:s_free", 10: "mma:S_issued", 11: "mma:p_ready", 12: "mma:PV_issued"}
for j in range(8, 11):
    ev = sorted((int(t[j, k]) - t0, names[k]) for k in names if int(t[j, k]) != 0)
    print(f"iter {j}: " + "  ".join(f"{n}@{c}" for c, n in ev))
print("period (sm:s_full):", [int(t[j + 1, 0] - t[j, 0]) for j in range(4, 14)])
