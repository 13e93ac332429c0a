"""LayerNorm folded into the consumer GEMM (vx_row_stats + vx_gemm_lnfold_bf16; engine switch VX_LN_FOLD=1)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("rows,C", [(4096, 320), (1000, 640), (513, 1280), (300, 128), (64, 2048)])
def test_row_stats(rows, C):
    from vexpress_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    x = (torch.randn(rows, C, device="cuda", generator=g) * 3 + 1.5).bfloat16()
    st = ops.row_stats(x)
    xf = x.float()
    torch.testing.assert_close(st[:, 0], xf.mean(1), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(st[:, 1], (xf.var(1, unbiased=False) + 1e-5).rsqrt(), atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("M,K,N,geglu,residual", [(4096, 320, 960, False, False), (2048, 1280, 1280, False, True),
                                                  (1024, 320, 2560, True, False), (8192, 1280, 10240, True, False),
                                                  (300, 640, 640, False, False)])
def test_gemm_lnfold_matches_layernorm_then_linear(M, K, N, geglu, residual):
    from vexpress_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = (torch.randn(M, K, device="cuda", generator=g) * 2 + 0.7).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(K, device="cuda", generator=g)
    beta = 0.1 * torch.randn(K, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if residual else None
    wf, cs, bf = ops.fold_layernorm(w, b, gamma, beta, geglu=geglu)
    out = ops.gemm_lnfold(x, wf, ops.row_stats(x), cs, bf, residual=res, geglu=geglu)
    ref = F.layer_norm(x.float(), (K,), gamma, beta, 1e-5) @ w.float().t() + b
    if geglu:
        h, gate = ref.chunk(2, dim=-1)
        ref = h * F.gelu(gate)
    if residual:
        ref = ref + res.float()
    err = _rel(out, ref)
    print(f"lnfold M={M} K={K} N={N} geglu={geglu} rel={err:.3e}")
    assert err < 6e-3, err


@pytest.mark.parametrize("M,K,N,geglu,residual,pe", [(4096, 320, 960, False, False, False), (4096, 320, 320, False, False, False),
                                                     (2048, 320, 2560, True, False, False), (4096, 320, 960, False, False, True),
                                                     (1000, 320, 960, False, True, False), (384, 320, 320, False, False, False),
                                                     (128, 320, 960, False, False, False), (100, 64, 64, False, False, False),
                                                     (2048, 512, 1024, False, True, False), (3072, 128, 512, True, False, False),
                                                     (40960, 320, 960, False, False, False), (20000, 320, 2560, True, False, False)])
def test_gemm_ln_matches_layernorm_then_linear(M, K, N, geglu, residual, pe):
    """vx_gemm_ln_bf16: LayerNorm -> Linear with the row tile resident in shared memory and the statistics computed in the
    kernel; against fp32 torch and, bit for bit run to run, against itself (row tiles walk all their column tiles: one, two
    and many row-tile groups per CTA pair, odd row-tile counts, ragged last tile, the single-CTA path at M <= 128)."""
    from vexpress_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = (torch.randn(M, K, device="cuda", generator=g) * 2 + 0.7).bfloat16()
    x[::7] *= 8.0           # rows of very different scale: the statistics are per row
    x[3::11] += 30.0        # large mean against the spread: E[x^2] - mean^2 would lose bits here, the two-pass form does not
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(K, device="cuda", generator=g)
    beta = 0.1 * torch.randn(K, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if residual else None
    f_, rows = 16, 64
    bias2 = torch.randn((M + rows - 1) // rows, N, device="cuda", generator=g) if pe else None
    wf, cs, bf = ops.fold_layernorm(w, b, gamma, beta, geglu=geglu)
    out = ops.gemm_ln(x, wf, cs, bf, 1e-5, bias2=bias2, bias2_div=rows, residual=res, geglu=geglu)
    ref = F.layer_norm(x.float(), (K,), gamma, beta, 1e-5) @ w.float().t() + b
    if pe:
        ref = ref + bias2.repeat_interleave(rows, 0)[:M]
    if geglu:
        h, gate = ref.chunk(2, dim=-1)
        ref = h * F.gelu(gate)
    if residual:
        ref = ref + res.float()
    err = _rel(out, ref)
    # the two-kernel path on the same data, for scale
    n = ops.layernorm(x, gamma, beta)
    if geglu:
        wg, bg, _ = ops.pack_geglu(w, b)
        two = ops.gemm(n, wg, bg, geglu=True)
    else:
        two = ops.gemm(n, w, b, residual=res)
        if pe:
            two = None
    e2 = _rel(two, ref) if two is not None else float("nan")
    print(f"gemm_ln M={M} K={K} N={N} geglu={geglu} res={residual} pe={pe}: rel={err:.3e} (layernorm kernel + gemm: {e2:.3e})")
    assert err < 6e-3, err
    for _ in range(3):
        torch.empty(1 << 22, device="cuda").normal_()
        again = ops.gemm_ln(x, wf, cs, bf, 1e-5, bias2=bias2, bias2_div=rows, residual=res, geglu=geglu)
        assert torch.equal(again, out)


@pytest.mark.parametrize("M,K,C,N,geglu,residual", [(4096, 320, 320, 960, False, True), (131072, 320, 320, 320, False, True),
                                                    (2048, 1280, 1280, 3840, False, True), (8192, 5120, 1280, 10240, True, True),
                                                    (1000, 640, 640, 1920, False, False), (300, 64, 64, 192, False, True),
                                                    (32768, 640, 640, 5120, True, True), (128 * 75, 256, 128, 256, False, True)])
def test_layernorm_handover_between_gemms(M, K, C, N, geglu, residual):
    """Producer GEMM [M,K]x[C,K] (+bias, +residual) with row sums -> consumer GEMM LayerNorm(h) @ W^T from the partial
    sums, against (a) the producer's plain twin: identical output bits, row sums equal to fp32 sums of those bits within
    fp32 rounding, and (b) LayerNorm kernel + GEMM on the same h / torch fp32."""
    from vexpress_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    wp = (torch.randn(C, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bp = torch.randn(C, device="cuda", generator=g)
    res = (torch.randn(M, C, device="cuda", generator=g) * 2 + 0.5).bfloat16() if residual else None
    h_plain = ops.gemm(a, wp, bp, residual=res)
    h, parts, n = ops.gemm_rowsums(a, wp, bp, residual=res)
    assert torch.equal(h, h_plain) and 2 <= n <= ops.rowsum_slots(C) == parts.shape[0]
    hf = h.float()
    s1, s2 = parts[:n, :, 0].sum(0), parts[:n, :, 1].sum(0)
    e1 = ((s1 - hf.sum(1)).abs() / hf.abs().sum(1).clamp_min(1e-6)).max().item()
    e2 = ((s2 - (hf * hf).sum(1)).abs() / (hf * hf).sum(1).clamp_min(1e-6)).max().item()
    w = (torch.randn(N, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    wf, cs, bf = ops.fold_layernorm(w, b, gamma, beta, geglu=geglu)
    out = ops.gemm_lnparts(h, wf, parts, n, cs, bf, 1e-5, geglu=geglu)
    ref = F.layer_norm(hf, (C,), gamma, beta, 1e-5) @ w.float().t() + b
    if geglu:
        wg, bg, _ = ops.pack_geglu(w, b)
        two = ops.gemm(ops.layernorm(h, gamma, beta, 1e-5), wg, bg, geglu=True)
        v, gate = ref.chunk(2, dim=-1)
        ref = v * F.gelu(gate)
    else:
        two = ops.gemm(ops.layernorm(h, gamma, beta, 1e-5), w, b)
    err, err2 = _rel(out, ref), _rel(two, ref)
    print(f"handover M={M} K={K} C={C} N={N} geglu={geglu}: {n} partials, row sums {e1:.1e} / {e2:.1e}; "
          f"rel={err:.3e} (layernorm kernel + gemm: {err2:.3e})")
    assert e1 < 1e-5 and e2 < 1e-5
    assert err < 6e-3 and err < 1.5 * err2 + 1e-4, (err, err2)


def test_unet_with_ln_fold_vs_oracle(golden_dir, monkeypatch):
    """Whole small UNet with VX_LN_FOLD=1 and with the statistics hand-over (VX_LN_FUSE=1) against the fp32 oracle on the same bf16-rounded weights: same bound as the
    default LayerNorm-kernel path (two bf16 evaluation orders of the same network sit ~sqrt(2) x 1.6e-2 apart from each
    other, so they are each compared with fp32, not with one another)."""
    from oracle import vx_oracle as O
    from test_unet_gpu import build_product
    cfg = O.small_cfg()
    sd = O.synth_state_dict(O.unet_param_shapes(cfg), 1234)
    lat, kps, audio, banks = O.synth_inputs(cfg, 4, 16, 16, True, 42)
    x = lat.repeat(2, 1, 1, 1, 1)
    enc = audio.reshape(-1, 5, cfg["cross_attention_dim"])
    r = lambda t: t.bfloat16().float()
    with torch.no_grad():
        ref = O.unet_forward({k: r(v) for k, v in sd.items()}, cfg, r(x), 499, r(enc), r(kps), [r(b) for b in banks], 0.95, 3.0)
    errs = {}
    for mode, (fold, fuse) in dict(kernel=("0", "0"), fold=("1", "0"), handover=("0", "1")).items():
        monkeypatch.setenv("VX_LN_FOLD", fold)
        monkeypatch.setenv("VX_LN_FUSE", fuse)
        model, _ = build_product(cfg, sd, [b[1:] for b in banks], 0.95, 3.0)
        eng = model.engine()
        assert (eng.ln_fold, eng.ln_fuse) == (fold == "1", fuse == "1")
        out = model(x.cuda().bfloat16(), 499, enc.cuda().bfloat16(), kps_features=kps.cuda().bfloat16(), return_dict=False)[0]
        errs[mode] = _rel(out.cpu(), ref)
    print(f"unet vs oracle: LayerNorm kernel {errs['kernel']:.3e}, ln-fold {errs['fold']:.3e}, statistics hand-over {errs['handover']:.3e}")
    for mode in ("fold", "handover"):
        assert errs[mode] < 3e-2 and errs[mode] < 1.5 * errs["kernel"], errs
