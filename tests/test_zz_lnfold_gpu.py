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


def test_unet_with_ln_fold_matches_default(golden_dir, monkeypatch):
    """Whole small UNet: VX_LN_FOLD=1 against the default LayerNorm-kernel path (same weights, same inputs)."""
    from oracle import vx_oracle as O
    from vexpress_b200.modules import ReferenceAttentionControl, UNet3DConditionModel
    cfg = O.small_cfg()
    sd = O.synth_state_dict(O.unet_param_shapes(cfg), 1234)
    lat, kps, audio, banks = O.synth_inputs(cfg, 4, 16, 16, True, 42)

    def build():
        m = UNet3DConditionModel(
            block_out_channels=cfg["block_out_channels"], cross_attention_dim=cfg["cross_attention_dim"],
            use_inflated_groupnorm=True, use_motion_module=True, motion_module_mid_block=True,
            motion_module_type="Vanilla",
            motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                                      attention_block_types=["Temporal_Self", "Temporal_Self"],
                                      temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                                      temporal_attention_dim_div=1))
        m.load_state_dict(sd, strict=True)
        m = m.to(device="cuda", dtype=torch.bfloat16)
        r = ReferenceAttentionControl(m, do_classifier_free_guidance=True, mode="read", fusion_blocks="full",
                                      reference_attention_weight=0.95, audio_attention_weight=3.0)
        r.update(type("W", (), {"banks": [b[1:].cuda() for b in banks]})(), True, dtype=torch.bfloat16)
        return m
    x = lat.repeat(2, 1, 1, 1, 1).cuda().bfloat16()
    enc = audio.reshape(-1, 5, cfg["cross_attention_dim"]).cuda().bfloat16()
    k = kps.cuda().bfloat16()
    base = build()(x, 499, enc, kps_features=k, return_dict=False)[0]
    monkeypatch.setenv("VX_LN_FOLD", "1")
    fold = build()(x, 499, enc, kps_features=k, return_dict=False)[0]
    err = _rel(fold, base)
    print(f"unet ln-fold vs default rel={err:.3e}")
    assert err < 2e-2, err
