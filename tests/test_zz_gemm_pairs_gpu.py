"""Corner cases of the CTA-pair (cta_group::2) GEMM / conv path that the UNet / VAE shapes never hit: an odd number of
row tiles (the last pair's second CTA works on an out-of-range tile), M not a multiple of 128, residual + bias2 on pairs."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("M,N,K", [(384, 256, 1024), (300, 640, 1280), (128 * 75, 320, 1280), (1000, 1280, 5120)])
def test_pair_gemm_odd_row_tiles(M, N, K):
    from vexpress_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    out = ops.gemm(a, w, bias, residual=res)
    ref = a.float() @ w.float().t() + bias + res.float()
    err = _rel(out, ref)
    print(f"pair gemm {M}x{N}x{K} rel={err:.3e}")
    assert err < 5e-3, err


@pytest.mark.parametrize("NB,H,W,C,Cout", [(3, 16, 24, 128, 128), (5, 8, 16, 256, 320), (1, 48, 128, 64, 64)])
def test_pair_conv_odd_row_tiles(NB, H, W, C, Cout):
    import torch.nn.functional as F
    from vexpress_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(NB * H + C)
    x = torch.randn(NB, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    out = ops.conv3x3(x, ops.pack_conv3x3_weight(w), b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    err = _rel(out, ref)
    print(f"pair conv NB={NB} {H}x{W} C={C}->{Cout} rel={err:.3e}")
    assert err < 5e-3, err


# ----------------------------------------------------------------------------- W multicast (VX_GEMM_MC=1)
@pytest.fixture
def w_multicast():
    """Switch the 1-CTA GEMM / conv launches to clusters of two CTAs that multicast the W tile halves to each other
    (GemmArgs::mc).  The library reads its switches once; the bring-up hook re-reads them."""
    from vexpress_b200 import _ffi
    old = os.environ.get("VX_GEMM_MC")

    def switch(on):
        if on:
            os.environ["VX_GEMM_MC"] = "1"
        else:
            os.environ.pop("VX_GEMM_MC", None)
        _ffi.lib().vx_gemm_reload_env()
    yield switch
    if old is None:
        os.environ.pop("VX_GEMM_MC", None)
    else:
        os.environ["VX_GEMM_MC"] = old
    _ffi.lib().vx_gemm_reload_env()


@pytest.mark.parametrize("M,N,K,res,geglu", [
    (256, 64, 64, False, False), (384, 256, 320, True, False), (300, 640, 640, True, False), (128 * 75, 320, 320, True, False),
    (131072, 320, 320, True, False), (131072, 960, 320, False, False), (32768, 640, 640, True, False), (32768, 1920, 640, False, False),
    (4096, 2560, 320, False, True), (32768, 5120, 640, False, True), (1000, 96, 200, False, False), (128, 320, 320, True, False)])
def test_gemm_w_multicast(w_multicast, M, N, K, res, geglu):
    """Same MMAs in the same order on the same operand bytes: the multicast launch must reproduce the plain launch bit for
    bit (and both sit at the usual distance from fp32), including odd row-tile counts (the cluster's second CTA then works
    on an out-of-range tile), M % 128 != 0, residual prefetch and the GEGLU epilogue."""
    from vexpress_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
    if geglu:
        wp, bp, _ = ops.pack_geglu(w, bias)
        run = lambda: ops.gemm(a, wp, bp, geglu=True)
        h, gate = (a.float() @ w.float().t() + bias).chunk(2, -1)
        ref = h * torch.nn.functional.gelu(gate)
    else:
        run = lambda: ops.gemm(a, w, bias, residual=r)
        ref = a.float() @ w.float().t() + bias + (r.float() if res else 0)
    w_multicast(False)
    plain = run().clone()
    w_multicast(True)
    outs = [run().clone() for _ in range(3)]
    torch.cuda.synchronize()
    err = _rel(outs[0], ref)
    same = all(torch.equal(o, plain) for o in outs)
    print(f"mc gemm {M}x{N}x{K} res={res} geglu={geglu} rel={err:.3e} identical to the plain launch: {same}")
    assert err < 5e-3 and same


@pytest.mark.parametrize("NB,H,W,C,Cout", [(2, 16, 16, 64, 64), (3, 16, 24, 64, 96), (1, 48, 128, 64, 64), (5, 8, 8, 64, 32)])
def test_conv_w_multicast(w_multicast, NB, H, W, C, Cout):
    """3x3 convs whose K loop is too short for the pair kernel (C = 64) take the multicast launch too."""
    import torch.nn.functional as F
    from vexpress_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(NB * H + C)
    x = torch.randn(NB, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    wp = ops.pack_conv3x3_weight(w)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    w_multicast(False)
    plain = ops.conv3x3(x, wp, b).clone()
    plain_s2 = ops.conv3x3_s2(x, wp, b).clone() if H % 2 == 0 and W % 2 == 0 else None
    w_multicast(True)
    out = ops.conv3x3(x, wp, b)
    out_s2 = ops.conv3x3_s2(x, wp, b) if plain_s2 is not None else None
    torch.cuda.synchronize()
    err = _rel(out, ref)
    print(f"mc conv NB={NB} {H}x{W} C={C}->{Cout} rel={err:.3e} identical: {torch.equal(out, plain)}")
    assert err < 5e-3 and torch.equal(out, plain)
    assert plain_s2 is None or torch.equal(out_s2, plain_s2)
