"""Attention, normalisation and misc kernels against torch fp32 on the same bf16-rounded inputs (GPU)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops():
    from vexpress_b200 import _ffi, ops
    _ffi.require_sm100()
    return ops


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("B,N,heads,hd,kv_div,Nk", [(2, 256, 8, 160, 1, 256), (2, 1024, 8, 80, 1, 1024),
                                                    (2, 4096, 8, 40, 1, 4096), (4, 64, 8, 160, 1, 64),
                                                    (4, 1024, 8, 80, 2, 1024), (4, 256, 8, 16, 2, 256),
                                                    (2, 128, 2, 32, 1, 384), (3, 64, 8, 8, 1, 64),
                                                    (8, 4, 8, 32, 1, 4), (4, 36, 8, 16, 2, 36),
                                                    # BASELINE configs[4] (768x768: 96x96 latents): N = 9216 / 2304 / 576 / 144
                                                    (2, 9216, 8, 40, 1, 9216), (4, 9216, 8, 40, 2, 9216),
                                                    (2, 2304, 8, 80, 1, 2304), (2, 576, 8, 160, 1, 576),
                                                    (4, 144, 8, 160, 1, 144), (4, 144, 8, 160, 2, 144),
                                                    # head dims of the 8-softmax-warp kernel (hd <= 64), incl. padded ones
                                                    (2, 128, 4, 64, 1, 192), (1, 256, 2, 48, 1, 128), (2, 128, 8, 8, 1, 64),
                                                    (3, 384, 8, 40, 3, 320), (2, 128, 8, 24, 1, 64)])
def test_flash_attention(ops, B, N, heads, hd, kv_div, Nk):
    g = _gen(B * N + hd)
    C = heads * hd
    Bkv = (B + kv_div - 1) // kv_div
    qkv = torch.randn(B * N, 3 * C, device="cuda", generator=g).bfloat16()
    q = qkv[:, :C]
    if kv_div == 1 and Nk == N:
        k, v = qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:
        kv = (1.5 * torch.randn(Bkv * Nk, 2 * C, device="cuda", generator=g)).bfloat16()
        k, v = kv[:, :C], kv[:, C:]
    out = ops.flash_attention(q, k, v, heads, N, Nk, kv_div)
    qf = q.float().view(B, N, heads, hd).transpose(1, 2)
    kf = k.float().reshape(Bkv, Nk, heads, hd).transpose(1, 2).repeat_interleave(kv_div, 0)[:B]
    vf = v.float().reshape(Bkv, Nk, heads, hd).transpose(1, 2).repeat_interleave(kv_div, 0)[:B]
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * N, C)
    torch.cuda.synchronize()
    err = _rel(out, ref)
    print(f"flash B={B} N={N} hd={hd} kv_div={kv_div} Nk={Nk} rel={err:.3e} nan={torch.isnan(out.float()).any().item()}")
    # bf16 P and bf16 output put a correct kernel at 2.3e-3 on this data; the hd 80 ordering hazard of round 2 (S(t) issued
    # behind a P.V(t-2) that had not completed) sat at 5e-3 .. 8e-3 -- under the 1e-2 this assertion used to allow
    assert err < 4e-3, err


@pytest.mark.parametrize("B,N,heads,hd", [(8, 1024, 8, 80), (4, 1024, 8, 96), (2, 4096, 8, 40), (4, 1024, 8, 64), (2, 2048, 8, 128),
                                          (4, 256, 8, 160)])
def test_flash_attention_is_deterministic(ops, B, N, heads, hd):
    """Identical inputs -> identical bits, run after run, with unrelated kernels in between (a tensor-memory or ring race shows
    up as run-to-run differences long before it moves the error against fp32 past a tolerance)."""
    g = torch.Generator(device="cuda").manual_seed(B * N + hd)
    C = heads * hd
    qkv = torch.randn(B * N, 3 * C, device="cuda", generator=g).bfloat16()
    first = None
    for i in range(12):
        out = ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, N, N)
        if i % 2 == 0:
            torch.empty(1 << 24, device="cuda").normal_()
        if first is None:
            first = out.clone()
        else:
            assert torch.equal(out, first), f"run {i} differs from run 0 by {(out.float() - first.float()).abs().max().item():.3e}"


@pytest.mark.parametrize("B,N,heads,hd", [(2, 128, 4, 40), (2, 256, 4, 40), (2, 1024, 4, 40), (2, 1024, 4, 64), (2, 128, 4, 80),
                                          (2, 1024, 4, 80), (2, 1024, 2, 128)])
def test_flash_attention_with_a_late_mma_warp(ops, B, N, heads, hd, monkeypatch):
    """Fault injection (bring-up switch VX_FA3_DBG bit 5): the MMA warp idles 3000 clocks in front of every P.V, so the
    softmax warps run as far ahead of the tensor core as the protocol lets them.  Every wait in the kernel must still mean what
    it says -- round 2's first v3 read O after a parity wait that had silently skipped a phase in exactly this situation
    (which the K/V-load-bound hd 80 shape produced on its own)."""
    from vexpress_b200 import _ffi
    g = torch.Generator(device="cuda").manual_seed(N + hd)
    C = heads * hd
    qkv = torch.randn(B * N, 3 * C, device="cuda", generator=g).bfloat16()
    ref = F.scaled_dot_product_attention(*[qkv[:, i * C:(i + 1) * C].float().view(B, N, heads, hd).transpose(1, 2) for i in range(3)])
    ref = ref.transpose(1, 2).reshape(B * N, C)
    errs = {}
    try:
        for dbg in ("0", "32", "34"):      # 34: additionally every S(t) completes before anything else is issued
            monkeypatch.setenv("VX_FA3_DBG", dbg)
            _ffi.lib().vx_flash_reload_env()
            errs[dbg] = _rel(ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, N, N), ref)
    finally:
        monkeypatch.delenv("VX_FA3_DBG")
        _ffi.lib().vx_flash_reload_env()
    print(f"flash hd={hd} N={N}: rel err normal {errs['0']:.3e}, MMA warp delayed {errs['32']:.3e}, delayed + serialised {errs['34']:.3e}")
    assert max(errs.values()) < 4e-3, errs


def test_flash_attention_zero_kv(ops):
    """CFG uncond half: all-zero bank -> K = V = 0 -> output exactly 0 (mutual_self_attention.py:359)."""
    g = _gen(3)
    q = torch.randn(2 * 256, 320, device="cuda", generator=g).bfloat16()
    z = torch.zeros(256, 320, device="cuda", dtype=torch.bfloat16)
    out = ops.flash_attention(q, z, z, 8, 256, 256, kv_div=2)
    assert torch.count_nonzero(out).item() == 0


@pytest.mark.parametrize("b,f,HW,heads,hd", [(2, 16, 64, 8, 40), (1, 4, 256, 8, 160), (2, 16, 16, 8, 8), (1, 32, 32, 8, 80)])
def test_temporal_attention(ops, b, f, HW, heads, hd):
    g = _gen(f + hd)
    C = heads * hd
    qkv = torch.randn(b * f * HW, 3 * C, device="cuda", generator=g).bfloat16()
    out = ops.temporal_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], b, f, HW, heads)
    x = qkv.float().view(b, f, HW, 3, heads, hd).permute(3, 0, 2, 4, 1, 5)      # (3, b, hw, heads, f, hd)
    ref = F.scaled_dot_product_attention(x[0], x[1], x[2])                      # (b, hw, heads, f, hd)
    ref = ref.permute(0, 3, 1, 2, 4).reshape(b * f * HW, C)
    err = _rel(out, ref)
    print(f"temporal b={b} f={f} HW={HW} hd={hd} rel={err:.3e}")
    assert err < 5e-3


@pytest.mark.parametrize("frames,N,heads,hd", [(4, 256, 8, 40), (2, 64, 8, 160), (3, 128, 8, 8)])
def test_smallkv_attention(ops, frames, N, heads, hd):
    g = _gen(N + hd)
    C = heads * hd
    q = torch.randn(frames * N, C, device="cuda", generator=g).bfloat16()
    kv = torch.randn(frames * 5, 2 * C, device="cuda", generator=g).bfloat16()
    out = ops.smallkv_attention(q, kv[:, :C], kv[:, C:], N, heads, 5)
    qf = q.float().view(frames, N, heads, hd).transpose(1, 2)
    kf = kv[:, :C].float().reshape(frames, 5, heads, hd).transpose(1, 2)
    vf = kv[:, C:].float().reshape(frames, 5, heads, hd).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(frames * N, C)
    err = _rel(out, ref)
    print(f"smallkv frames={frames} N={N} hd={hd} rel={err:.3e}")
    assert err < 5e-3


@pytest.mark.parametrize("NB,HW,C1,C2,silu,eps", [(4, 4096, 320, 0, True, 1e-5), (3, 256, 1280, 640, True, 1e-5),
                                                  (2, 64, 1280, 1280, False, 1e-6), (5, 1024, 64, 0, False, 1e-6),
                                                  (2, 256, 640, 320, True, 1e-5), (1, 65536, 128, 0, True, 1e-6),
                                                  (3, 9216, 320, 0, True, 1e-5), (2, 2304, 640, 320, True, 1e-5),
                                                  (2, 144, 1280, 1280, True, 1e-5), (1, 147456, 128, 0, True, 1e-6)])
def test_groupnorm(ops, NB, HW, C1, C2, silu, eps):
    g = _gen(C1 + C2 + HW)
    x1 = (torch.randn(NB * HW, C1, device="cuda", generator=g) * 2 + 0.7).bfloat16()
    x2 = (torch.randn(NB * HW, C2, device="cuda", generator=g) - 0.3).bfloat16() if C2 else None
    C = C1 + C2
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    out = ops.groupnorm(x1, NB, HW, gamma, beta, eps, silu, x2=x2)
    x = x1 if x2 is None else torch.cat([x1, x2], 1)
    ref = F.group_norm(x.float().view(NB, HW, C).transpose(1, 2), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.transpose(1, 2).reshape(NB * HW, C)
    err = _rel(out, ref)
    print(f"groupnorm NB={NB} HW={HW} C={C1}+{C2} rel={err:.3e}")
    assert err < 4e-3


@pytest.mark.parametrize("NB,HW,C1,C2,silu", [(32, 64, 1280, 0, True), (32, 64, 1280, 1280, True), (32, 256, 1280, 0, False),
                                              (32, 256, 1280, 640, True), (8, 64, 1280, 0, True), (2, 256, 640, 640, True),
                                              (4, 64, 320, 0, False), (3, 16, 256, 0, True), (32, 1024, 640, 0, True),
                                              (5, 36, 64, 32, True)])
def test_groupnorm_cluster_resident(ops, NB, HW, C1, C2, silu):
    """GroupNorm with the frame resident in a thread-block cluster's shared memory (one pass over HBM, partial statistics
    through distributed shared memory) vs torch fp32 and vs the one-launch rendezvous kernel; repeated launches must give
    identical bits (fixed merge order).  Frames that do not fit one wave of clusters fall back to the rendezvous kernel
    (the 16x16 x 32-frame and 32x32 cases: identical bits to it)."""
    g = _gen(NB + HW + C1)
    x1 = (torch.randn(NB * HW, C1, device="cuda", generator=g) * 2 + 0.3).bfloat16()
    x2 = (torch.randn(NB * HW, C2, device="cuda", generator=g) - 0.5).bfloat16() if C2 else None
    C = C1 + C2
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    old = ops._GN_CLUSTER
    try:
        ops._GN_CLUSTER = False
        base = ops.groupnorm(x1, NB, HW, gamma, beta, 1e-5, silu, x2=x2).clone()
        ops._GN_CLUSTER = True
        outs = [ops.groupnorm(x1, NB, HW, gamma, beta, 1e-5, silu, x2=x2).clone() for _ in range(3)]
    finally:
        ops._GN_CLUSTER = old
    x = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], 1)
    ref = F.group_norm(x.view(NB, HW, C).transpose(1, 2), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.transpose(1, 2).reshape(NB * HW, C)
    err, err_base = _rel(outs[0], ref), _rel(base, ref)
    print(f"groupnorm cluster NB={NB} HW={HW} C={C1}+{C2} rel={err:.3e} (rendezvous kernel {err_base:.3e}), "
          f"vs rendezvous kernel {_rel(outs[0], base.float()):.2e}")
    assert err < 4e-3 and torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("rows,C,with_pe", [(4096, 320, False), (1000, 1280, True), (512, 640, True), (777, 64, False),
                                            (2 * 9216, 320, False), (2 * 2304 + 3, 640, False), (4 * 144, 1280, True),
                                            (4096 + 5, 320, True), (131072, 320, True)])
def test_layernorm(ops, rows, C, with_pe):
    g = _gen(rows + C)
    x = (torch.randn(rows, C, device="cuda", generator=g) * 3 + 1).bfloat16()
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    pe = torch.randn(16, C, device="cuda", generator=g) if with_pe else None
    rpf = 25
    out = ops.layernorm(x, gamma, beta, 1e-5, pe=pe, rows_per_frame=rpf)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    if with_pe:
        fr = (torch.arange(rows, device="cuda") // rpf) % 16
        ref = ref + pe[fr]
    err = _rel(out, ref)
    print(f"layernorm rows={rows} C={C} pe={with_pe} rel={err:.3e}")
    assert err < 4e-3


def test_geglu(ops):
    g = _gen(5)
    x = torch.randn(1000, 2 * 1280, device="cuda", generator=g).bfloat16()
    out = ops.geglu(x)
    h, gate = x.float().chunk(2, -1)
    assert _rel(out, h * F.gelu(gate)) < 4e-3


def test_conv_in_out_im2col_upsample(ops):
    g = _gen(9)
    NB, H, W, Cout = 6, 32, 32, 320
    x = torch.randn(2, 4, 3, H, W, device="cuda", generator=g).bfloat16()                   # (b, c, f, h, w)
    w = torch.randn(Cout, 4, 3, 3, device="cuda", generator=g) / 6
    bias = torch.randn(Cout, device="cuda", generator=g)
    kps = torch.randn(10 * H * W, Cout, device="cuda", generator=g).bfloat16()
    frames = torch.tensor([7, 1, 3, 9, 0, 2], device="cuda", dtype=torch.int32)
    xin = x.permute(0, 2, 1, 3, 4).reshape(NB, 4, H, W)                                     # strided view, planes contiguous
    out = ops.conv_in(xin, w.reshape(Cout, 36).t().contiguous(), bias, Cout, addend=kps, add_frame=frames)
    ref = F.conv2d(xin.float(), w, bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    ref = ref + kps.float().view(10, H * W, Cout)[frames.long()].reshape(-1, Cout)
    assert _rel(out, ref) < 4e-3

    # conv_out: NHWC -> frame-major planar (n, co, h, w)
    C = 320
    y = torch.randn(NB * H * W, C, device="cuda", generator=g).bfloat16()
    w2 = torch.randn(4, C, 3, 3, device="cuda", generator=g) / 50
    b2 = torch.randn(4, device="cuda", generator=g)
    o = torch.empty(NB, 4, H, W, device="cuda", dtype=torch.bfloat16)
    ops.conv_out(y, NB, H, W, w2.permute(0, 2, 3, 1).reshape(4, 9, C).contiguous(), b2, o)
    ref2 = F.conv2d(y.float().view(NB, H, W, C).permute(0, 3, 1, 2), w2, b2, padding=1)
    assert _rel(o, ref2) < 4e-3
    o32 = torch.empty(NB, 4, H, W, device="cuda", dtype=torch.float32)
    ops.conv_out(y, NB, H, W, w2.permute(0, 2, 3, 1).reshape(4, 9, C).contiguous(), b2, o32, post=True)
    assert _rel(o32, (ref2 / 2 + 0.5).clamp(0, 1)) < 1e-3

    # im2col stride 2 + gemm == conv stride 2
    xs = torch.randn(NB, H, W, 64, device="cuda", generator=g).bfloat16()
    ws = (torch.randn(128, 64, 3, 3, device="cuda", generator=g) / 24).bfloat16()
    col = ops.im2col_s2(xs, NB, H, W)
    o3 = ops.gemm(col, ops.pack_conv3x3_weight(ws))
    ref3 = F.conv2d(xs.float().permute(0, 3, 1, 2), ws.float(), stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, 128)
    assert _rel(o3, ref3) < 5e-3

    up = ops.upsample2x(xs, NB, H, W)
    refu = F.interpolate(xs.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).reshape(-1, 64)
    assert torch.equal(up.float(), refu)


def test_skinny_and_timestep(ops):
    g = _gen(11)
    t = torch.tensor([999.0, 499.0], device="cuda")
    emb = ops.timestep_embed(t, 320)
    half = 160
    fr = torch.exp(-math.log(10000) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    ref = torch.cat([torch.cos(t[:, None] * fr), torch.sin(t[:, None] * fr)], -1)
    assert (emb - ref).abs().max().item() < 1e-2          # bf16-rounded like the reference's cast
    w = (torch.randn(1280, 320, device="cuda", generator=g) / 18).bfloat16()
    b = torch.randn(1280, device="cuda", generator=g)
    y = ops.skinny_linear(emb, w, b, act_in=False, act_out=True)
    refy = F.silu(emb @ w.float().t() + b)
    assert _rel(y, refy) < 1e-4
    y2 = ops.skinny_linear(y, (torch.randn(640, 1280, device="cuda", generator=g) / 36).bfloat16(), None, act_in=True)
    assert y2.shape == (2, 640) and torch.isfinite(y2).all()


def test_cfg_overlap_ddim_matches_bf16_eager(ops):
    """Bit-exact against the reference's model-dtype (bf16) eager arithmetic (v_express_pipeline.py:548-572)."""
    g = _gen(13)
    f, hw, L, gs = 4, 64, 6, 3.5
    noise = torch.randn(2, f, 4, hw, device="cuda", generator=g).bfloat16()      # ((b f), 4, hw)
    win = torch.tensor([2, 3, 4, 5], device="cuda", dtype=torch.int32)
    count = torch.tensor([1, 1, 2, 2, 1, 1], device="cuda", dtype=torch.int32)
    acc = torch.zeros(4, L, hw, device="cuda")
    prev = torch.randn(4, L, hw, device="cuda", generator=g).bfloat16()
    acc.copy_(prev.float())
    acc[:, :2] = 0
    acc[:, 4:] = 0
    ops.cfg_overlap_accumulate(noise, f, hw, L, True, win, count, gs, acc)
    u, c = noise[0].transpose(0, 1), noise[1].transpose(0, 1)   # (4, f, hw)
    npred = u + gs * (c - u)                                    # bf16 eager
    npred = npred / count[win.long()].to(torch.bfloat16)[None, :, None]
    want = torch.zeros(4, L, hw, device="cuda", dtype=torch.bfloat16)
    want[:, 2:4] = prev[:, 2:4]
    want[:, win.long()] += npred
    assert torch.equal(acc.bfloat16(), want) and torch.equal(acc, want.float())

    lat = torch.randn(4, L, hw, device="cuda", generator=g).bfloat16()
    a_t, a_p = torch.tensor(0.2423590), torch.tensor(0.3751530)
    x, v = lat.clone(), acc.bfloat16()
    x0 = (a_t ** 0.5) * x - ((1 - a_t) ** 0.5) * v
    eps = (a_t ** 0.5) * v + ((1 - a_t) ** 0.5) * x
    ref = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps
    ops.ddim_step(lat, acc, float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(a_p ** 0.5), float((1 - a_p) ** 0.5))
    assert torch.equal(lat, ref)


@pytest.mark.parametrize("NB,H,W,C,Cout", [(2, 8, 8, 1280, 1280), (3, 16, 16, 128, 64), (2, 32, 32, 640, 640),
                                           (1, 64, 64, 512, 512), (1, 128, 128, 64, 64), (2, 256, 256, 64, 32),
                                           (5, 12, 12, 64, 96), (1, 8, 24, 128, 128)])
def test_upconv3x3_equals_upsample_then_conv(ops, NB, H, W, C, Cout):
    """Nearest-2x upsample folded into the 3x3 conv (four parity-class 2x2 convolutions) vs torch on the same bf16 inputs.
    Tolerance: one extra bf16 rounding of the pre-summed weights (2^-9 relative) on top of the output rounding."""
    g = _gen(NB * H + C + Cout)
    x = torch.randn(NB, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    out = ops.upconv3x3(x, ops.pack_upconv_weight(w), b)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert out.shape == ref.shape
    err = _rel(out, ref)
    print(f"upconv NB={NB} {H}x{W} C={C}->{Cout} rel={err:.3e}")
    assert err < 5e-3, err


@pytest.mark.parametrize("NB,HW,C1,C2,silu", [(32, 4096, 320, 0, True), (32, 1024, 640, 320, True), (32, 64, 1280, 1280, True),
                                              (16, 16384, 512, 0, False), (3, 9216, 320, 0, True), (1, 256, 64, 0, False)])
def test_groupnorm_one_launch_equals_two_kernel_pair(ops, NB, HW, C1, C2, silu):
    """The fused (per-frame rendezvous) GroupNorm merges the same partial statistics in the same order as the
    statistics + apply pair: bit-identical, launch after launch (the rendezvous counters recycle themselves)."""
    g = _gen(NB + HW + C1)
    x1 = (torch.randn(NB * HW, C1, device="cuda", generator=g) * 2 + 0.7).bfloat16()
    x2 = (torch.randn(NB * HW, C2, device="cuda", generator=g) - 0.3).bfloat16() if C2 else None
    C = C1 + C2
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    was = ops._GN_FUSED
    try:
        ops._GN_FUSED = False
        ref = ops.groupnorm(x1, NB, HW, gamma, beta, 1e-5, silu, x2=x2)
        ops._GN_FUSED = True
        outs = [ops.groupnorm(x1, NB, HW, gamma, beta, 1e-5, silu, x2=x2) for _ in range(3)]
        torch.cuda.synchronize()
    finally:
        ops._GN_FUSED = was
    for o in outs:
        assert torch.equal(o, ref)
