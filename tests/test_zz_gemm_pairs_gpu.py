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
