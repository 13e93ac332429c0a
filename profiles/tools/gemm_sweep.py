"""Time the dominant UNet / VAE GEMM and implicit-GEMM conv shapes under different VX_GEMM_* settings.
usage: python profiles/tools/gemm_sweep.py            (run on a B200; prints TFLOP/s per shape and setting)"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vexpress_b200 import _ffi, ops
torch.manual_seed(0)
dev = 'cuda'

def bf(*s): return (torch.randn(*s, device=dev) * 0.05).bfloat16()

def t_ms(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

cases = []
def add_gemm(M, K, N, residual=False):
    a, w, b = bf(M, K), bf(N, K), torch.randn(N, device=dev)
    r = bf(M, N) if residual else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    cases.append((f"gemm M={M} K={K} N={N}{' +res' if residual else ''}", 2.0 * M * K * N,
                  lambda: ops.gemm(a, w, b, residual=r, out=out)))
def add_conv(NB, H, W, C, Cout, residual=False):
    x, w, b = bf(NB, H, W, C), bf(Cout, 9 * C), torch.randn(Cout, device=dev)
    r = bf(NB * H * W, Cout) if residual else None
    out = torch.empty(NB * H * W, Cout, device=dev, dtype=torch.bfloat16)
    cases.append((f"conv NB={NB} {H}x{W} C={C} Cout={Cout}{' +res' if residual else ''}", 2.0 * NB * H * W * 9 * C * Cout,
                  lambda: ops.conv3x3(x, w, b, residual=r, out=out)))

def add_geglu(M, K, N):
    a, w, b = bf(M, K), bf(N, K), torch.randn(N, device=dev)
    wp, bp, _ = ops.pack_geglu(w, b)
    out = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
    cases.append((f"gemm M={M} K={K} N={N} geglu", 2.0 * M * K * N, lambda: ops.gemm(a, wp, bp, geglu=True, out=out)))
add_geglu(131072, 320, 2560); add_geglu(32768, 640, 5120); add_geglu(8192, 1280, 10240)
add_gemm(131072, 320, 960); add_gemm(131072, 320, 320, True); add_gemm(131072, 1280, 320, True)
add_gemm(32768, 640, 1920); add_gemm(32768, 640, 640, True); add_gemm(32768, 2560, 640, True)
add_gemm(8192, 1280, 3840); add_gemm(8192, 1280, 1280, True); add_gemm(8192, 5120, 1280, True)
add_conv(32, 64, 64, 320, 320); add_conv(32, 64, 64, 640, 320); add_conv(32, 32, 32, 640, 640)
add_conv(32, 32, 32, 1280, 640); add_conv(32, 16, 16, 1280, 1280); add_conv(32, 16, 16, 2560, 1280)
add_conv(16, 128, 128, 512, 512); add_conv(16, 256, 256, 256, 256); add_conv(16, 512, 512, 128, 128)

settings = [("cg1", {"VX_GEMM_CG": "1"}), ("auto", {}), ("mc", {"VX_GEMM_MC": "1"}), ("cg1 mc", {"VX_GEMM_CG": "1", "VX_GEMM_MC": "1"}), ("cg2", {"VX_GEMM_CG": "2"}), ("cg2 nbuf1", {"VX_GEMM_CG": "2", "VX_GEMM_NBUF": "1"}),
            ("cg2 bn128", {"VX_GEMM_CG": "2", "VX_GEMM_BN": "128"}),
            ("cg1 bn128", {"VX_GEMM_CG": "1", "VX_GEMM_BN": "128"}), ("bn128", {"VX_GEMM_BN": "128"}), ("bn64", {"VX_GEMM_BN": "64"}),
            ("bn160", {"VX_GEMM_BN": "160"}), ("bn256", {"VX_GEMM_BN": "256"}), ("bn320", {"VX_GEMM_BN": "320"}),
            ("nbuf1", {"VX_GEMM_NBUF": "1"}), ("st2", {"VX_GEMM_STAGES": "2"}), ("st4", {"VX_GEMM_STAGES": "4"})]
if len(sys.argv) > 1:
    settings = [s for s in settings if s[0] in sys.argv[1:]]
keys = ["VX_GEMM_CG", "VX_GEMM_NBUF", "VX_GEMM_BN", "VX_GEMM_STAGES", "VX_GEMM_MC"]
os.environ["VX_GEMM_VERBOSE"] = "1"
for name, flop, fn in cases: fn()
torch.cuda.synchronize()
os.environ.pop("VX_GEMM_VERBOSE")
print(f"{'shape':48s}" + "".join(f"{n:>12s}" for n, _ in settings))
for name, flop, fn in cases:
    row = f"{name:48s}"
    for _, env in settings:
        for k in keys: os.environ.pop(k, None)
        os.environ.update(env)
        _ffi.lib().vx_gemm_reload_env()          # the library reads its switches once; the sweep asks it to re-read them
        try:
            row += f"{flop / t_ms(fn) * 1e-9:12.0f}"
        except Exception as e:
            row += f"{'err':>12s}"
    print(row, flush=True)
