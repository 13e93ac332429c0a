import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import ops
torch.manual_seed(0)
B,N,heads,hd=32,4096,8,40
C=heads*hd
qkv=torch.randn(B*N,3*C,device='cuda').bfloat16()
def run(dbg):
    os.environ['VX_FA_DBG']=str(dbg)
    for _ in range(2): ops.flash_attention(qkv[:,:C],qkv[:,C:2*C],qkv[:,2*C:],heads,N,N)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.flash_attention(qkv[:,:C],qkv[:,C:2*C],qkv[:,2*C:],heads,N,N)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/5
print(f"flash level-0 self-attention (B=32, N=4096, 8 heads, hd=40): {run(0):8.3f} ms")
