"""3x3 implicit-GEMM conv: row-reuse operand staging (one A box of hbox + 2 image rows per (dx, channel block) feeding the
three dy taps; default) against the tap-by-tap staging (VX_CONV_RR=0), on the UNet / VAE shapes.   TFLOP/s per shape."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import _ffi, ops
torch.manual_seed(0)
dev = 'cuda'


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.05).bfloat16()


def t_ms(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


shapes = [(32, 64, 64, 320, 320, True), (32, 64, 64, 640, 320, False), (32, 64, 64, 960, 320, False), (32, 32, 32, 640, 640, True),
          (32, 32, 32, 1280, 640, False), (32, 32, 32, 1920, 640, False), (32, 16, 16, 1280, 1280, True), (32, 16, 16, 2560, 1280, False),
          (32, 8, 8, 1280, 1280, True), (16, 64, 64, 512, 512, True), (16, 128, 128, 512, 512, False)]
settings = [("row reuse", {}), ("rr bn128", {"VX_GEMM_BN": "128"}), ("rr bn64", {"VX_GEMM_BN": "64"}), ("tap by tap", {"VX_CONV_RR": "0"})]
print(f"{'conv shape':44s}" + "".join(f"{n:>14s}" for n, _ in settings))
for NB, H, W, C, Cout, res in shapes:
    x, w, b = bf(NB, H, W, C), bf(Cout, 9 * C), torch.randn(Cout, device=dev)
    r = bf(NB * H * W, Cout) if res else None
    out = torch.empty(NB * H * W, Cout, device=dev, dtype=torch.bfloat16)
    flop = 2.0 * NB * H * W * 9 * C * Cout
    row, outs = f"NB={NB} {H}x{W} C={C}->{Cout}{' +res' if res else ''}".ljust(44), []
    for name, env in settings:
        for k in ("VX_CONV_RR", "VX_GEMM_BN"):
            os.environ.pop(k, None)
        os.environ.update(env)
        _ffi.lib().vx_gemm_reload_env()
        try:
            ms = t_ms(lambda: ops.conv3x3(x, w, b, residual=r, out=out))
            outs.append(out.clone())
            row += f"{flop / ms * 1e-9:14.0f}"
        except Exception as e:
            row += f"{'err':>14s}"
    d = max((o.float() - outs[-1].float()).abs().max().item() for o in outs) if outs else 0.0
    print(row + f"   max |diff| between variants {d:.2e}", flush=True)
for k in ("VX_CONV_RR", "VX_GEMM_BN"):
    os.environ.pop(k, None)
_ffi.lib().vx_gemm_reload_env()
