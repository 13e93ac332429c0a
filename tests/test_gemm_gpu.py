"""tcgen05 GEMM / implicit-GEMM conv against torch fp32 on the same bf16-rounded inputs (GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops():
    from vexpress_b200 import _ffi, ops
    _ffi.require_sm100()
    return ops


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (256, 128, 128, 128), (512, 320, 320, 160),
                                      (4096, 1280, 1280, 256), (160, 640, 768, 0), (2048, 2560, 320, 0),
                                      (131072, 320, 320, 0), (300, 96, 200, 0)])
def test_gemm_plain(ops, M, N, K, bn):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    out = ops.gemm(a, w, block_n=bn)
    ref = a.float() @ w.float().t()
    torch.cuda.synchronize()
    err = _rel(out, ref)
    print(f"gemm {M}x{N}x{K} bn={bn} rel={err:.3e}")
    assert err < 5e-3, err          # bf16 output rounding ~ 2^-9 relative


def test_gemm_epilogue_and_splitk(ops):
    g = torch.Generator(device="cuda").manual_seed(7)
    M, K1, K2, N = 1024, 640, 320, 640
    a = torch.randn(M, K1, device="cuda", generator=g).bfloat16()
    a2 = torch.randn(M, K2, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K1 + K2, device="cuda", generator=g) / 30).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    bias2 = torch.randn(4, N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    out = ops.gemm(a, w, bias, a2=a2, bias2=bias2, bias2_div=256, scale=0.95, residual=res)
    ref = (torch.cat([a, a2], 1).float() @ w.float().t() + bias + bias2.repeat_interleave(256, 0)) * 0.95 + res.float()
    err = _rel(out, ref)
    print("gemm epilogue rel", err)
    assert err < 5e-3


@pytest.mark.parametrize("NB,H,W,C,Cout", [(2, 64, 64, 64, 64), (4, 32, 32, 128, 256), (4, 16, 16, 256, 320),
                                           (6, 8, 8, 128, 128), (32, 64, 64, 320, 320), (1, 256, 256, 128, 128),
                                           (2, 8, 8, 2560, 1280),
                                           # row-reuse staging (tile = 2 / 4 / 8 whole image rows of one frame)
                                           (3, 32, 32, 640, 640), (2, 16, 16, 1280, 1280), (1, 64, 64, 960, 320),
                                           (5, 16, 16, 256, 64), (2, 64, 64, 512, 512)])
def test_conv3x3(ops, NB, H, W, C, Cout):
    g = torch.Generator(device="cuda").manual_seed(NB * H + C)
    x = torch.randn(NB, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(NB * H * W, Cout, device="cuda", generator=g).bfloat16()
    out = ops.conv3x3(x, ops.pack_conv3x3_weight(w), bias)
    out_r = ops.conv3x3(x, ops.pack_conv3x3_weight(w), bias, residual=res, scale=0.5)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    err = _rel(out, ref)
    err_r = _rel(out_r, ref * 0.5 + res.float())
    print(f"conv {NB}x{H}x{W}x{C}->{Cout} rel={err:.3e} (+scale, residual: {err_r:.3e})")
    assert err < 5e-3 and err_r < 5e-3


@pytest.mark.parametrize("NB,H,W,C,Cout", [(2, 16, 16, 64, 64), (4, 32, 32, 128, 256), (32, 64, 64, 320, 320),
                                           (32, 32, 32, 640, 640), (32, 16, 16, 1280, 1280), (3, 8, 8, 128, 64),
                                           (1, 512, 512, 128, 128), (2, 256, 256, 256, 256), (2, 24, 40, 64, 96)])
@pytest.mark.parametrize("pad_lo", [1, 0])
def test_conv3x3_stride2(ops, NB, H, W, C, Cout, pad_lo):
    """Stride-2 3x3 conv through the TMA traversal stride (no im2col tensor) vs torch fp32 and vs the gathered path
    (im2col + GEMM: same K order, same MMAs -> identical bits).  pad_lo=1: nn.Conv2d(stride=2, padding=1), the UNet
    downsamplers (reference modules/resnet.py:93-120); pad_lo=0: F.pad(x, (0,1,0,1)) + padding 0 (VAE encoder)."""
    F = torch.nn.functional
    g = torch.Generator(device="cuda").manual_seed(NB * H + C + pad_lo)
    x = torch.randn(NB, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    wp = ops.pack_conv3x3_weight(w)
    out = ops.conv3x3_s2(x, wp, bias, pad_lo=pad_lo)
    xin = x.permute(0, 3, 1, 2).float()
    if pad_lo == 0:
        ref = F.conv2d(F.pad(xin, (0, 1, 0, 1)), w.float(), bias, stride=2)
    else:
        ref = F.conv2d(xin, w.float(), bias, stride=2, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    xt = x.view(NB * H * W, C)
    col = ops.im2col_s2(xt, NB, H, W) if pad_lo == 1 else ops.im2col3x3(xt, NB, H, W, stride=2, pad_lo=0)
    gathered = ops.gemm(col, wp, bias)
    err = _rel(out, ref)
    same = torch.equal(out, gathered)
    print(f"conv s2 pad_lo={pad_lo} {NB}x{H}x{W}x{C}->{Cout} rel={err:.3e} identical to im2col+gemm: {same} "
          f"(max diff {(out.float() - gathered.float()).abs().max().item():.2e})")
    assert out.shape == (NB * (H // 2) * (W // 2), Cout) and err < 5e-3
    assert _rel(out, gathered.float()) < 1e-3


@pytest.mark.parametrize("M,C", [(1024, 64), (4096, 320), (2048, 1280), (300, 128)])
def test_gemm_geglu_epilogue(ops, M, C):
    g = torch.Generator(device="cuda").manual_seed(M + C)
    x = torch.randn(M, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(8 * C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    b = torch.randn(8 * C, device="cuda", generator=g)
    wp, bp, bn = ops.pack_geglu(w, b)
    out = ops.gemm(x, wp, bp, geglu=True)
    h, gate = (x.float() @ w.float().t() + b).chunk(2, -1)
    ref = h * torch.nn.functional.gelu(gate)
    err = _rel(out, ref)
    print(f"geglu-gemm M={M} C={C} bn={bn} rel={err:.3e}")
    assert out.shape == (M, 4 * C) and err < 5e-3


def test_gemm_fp32_out(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(4096, 512, device="cuda", generator=g).bfloat16()
    k = torch.randn(4096, 512, device="cuda", generator=g).bfloat16()
    out = ops.gemm(q, k, scale=512 ** -0.5, out_f32=True)
    ref = (q.float() @ k.float().t()) * 512 ** -0.5
    assert out.dtype == torch.float32 and (out - ref).abs().max().item() < 1e-3
