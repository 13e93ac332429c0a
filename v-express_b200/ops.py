"""Operator-level host wrappers over the C ABI (torch tensors in, raw pointers across the boundary).

Activations are bf16, channels-last token matrices ``[rows, C]`` (rows = (b f) h w); weights are bf16
``[N, K]`` (nn.Linear layout; conv weights repacked to ``[Cout, (ky kx cin)]``); biases fp32.
All launches go to torch's current CUDA stream, are asynchronous and CUDA-graph capturable."""
import ctypes

import torch

from . import _ffi
from ._ffi import c_float, c_int, c_ll, check, ptr, stream_ptr

BF16 = torch.bfloat16


def _chk_bf16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == BF16 and t.stride(-1) == 1, (t.dtype, t.device, t.stride())


def gemm(a, w, bias=None, *, a2=None, bias2=None, bias2_div=1, scale=1.0, residual=None, out=None, block_n=0):
    """out = (concat(a, a2) @ w.T + bias + bias2[row // bias2_div]) * scale + residual."""
    _chk_bf16(a, w, a2, residual, out)
    M, K1 = a.shape
    K2 = 0 if a2 is None else a2.shape[1]
    N = w.shape[0]
    assert w.shape[1] == K1 + K2
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=BF16)
    check(_ffi.lib().vx_gemm_bf16(
        ptr(a), c_ll(a.stride(0)), c_int(K1), ptr(a2), c_ll(0 if a2 is None else a2.stride(0)), c_int(K2),
        ptr(w), c_ll(w.stride(0)), c_int(M), c_int(N), ptr(bias), ptr(bias2), c_int(bias2_div), c_float(scale),
        ptr(residual), c_ll(0 if residual is None else residual.stride(0)), ptr(out), c_ll(out.stride(0)),
        c_int(block_n), stream_ptr()), "vx_gemm_bf16")
    return out


def conv3x3(x, w, bias=None, *, bias2=None, bias2_div=1, scale=1.0, residual=None, out=None, block_n=0):
    """x: NHWC bf16 [NB,H,W,C]; w: [Cout, 9*C]; returns [NB*H*W, Cout] (= NHWC)."""
    _chk_bf16(x, w, residual, out)
    assert x.is_contiguous()
    NB, H, W, C = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == 9 * C
    if out is None:
        out = torch.empty((NB * H * W, Cout), device=x.device, dtype=BF16)
    check(_ffi.lib().vx_conv3x3_bf16(
        ptr(x), c_int(NB), c_int(H), c_int(W), c_int(C), ptr(w), c_int(Cout), ptr(bias), ptr(bias2),
        c_int(bias2_div), c_float(scale), ptr(residual), c_ll(0 if residual is None else residual.stride(0)),
        ptr(out), c_ll(out.stride(0)), c_int(block_n), stream_ptr()), "vx_conv3x3_bf16")
    return out


def pack_conv3x3_weight(w):
    """(Cout, Cin, 3, 3) -> [Cout, (ky kx cin)] bf16, the K order the implicit-GEMM producer walks."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous()


# ----------------------------------------------------------------------------- bring-up probes
def probe_umma(a_img, b_img, lboA, sboA, layA, lboB, sboB, layB, a_mn, b_mn, N, ksteps, a_step, b_step):
    out = torch.empty((128, N), device=a_img.device, dtype=torch.float32)
    check(_ffi.lib().vx_probe_umma(ptr(a_img), c_int(a_img.numel()), ptr(b_img), c_int(b_img.numel()),
                                   ctypes.c_uint(lboA), ctypes.c_uint(sboA), ctypes.c_uint(layA),
                                   ctypes.c_uint(lboB), ctypes.c_uint(sboB), ctypes.c_uint(layB), c_int(a_mn),
                                   c_int(b_mn), c_int(N), c_int(ksteps), c_int(a_step), c_int(b_step), ptr(out),
                                   stream_ptr()), "vx_probe_umma")
    return out


def probe_tma(base, dims, strides_bytes, box, swizzle, coords, nbytes):
    rank = len(dims)
    out = torch.empty(nbytes, device=base.device, dtype=torch.uint8)
    U64 = ctypes.c_ulonglong * 5
    U32 = ctypes.c_uint * 5
    I32 = ctypes.c_int * 5
    pad = lambda x, n: list(x) + [0] * (n - len(x))
    check(_ffi.lib().vx_probe_tma(ptr(base), c_int(rank), U64(*pad(dims, 5)), U64(*pad(strides_bytes, 5)),
                                  U32(*pad(box, 5)), c_int(swizzle), I32(*pad(coords, 5)), c_int(nbytes), ptr(out),
                                  stream_ptr()), "vx_probe_tma")
    return out
