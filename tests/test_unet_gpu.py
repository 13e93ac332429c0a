"""Whole denoising-UNet forward on the B200 kernels against the oracle / reference golden (GPU).

Tolerance (stated per north_star): the product computes in bf16 with fp32 accumulation; against the fp32
oracle evaluated on the SAME bf16-rounded weights and inputs the relative L2 error of every block-boundary
activation must stay below 2e-2 and of the final noise prediction below 3e-2 (bf16 has 8 mantissa bits:
one rounding is 2^-9 ~ 2e-3 relative; ~60 sequential roundings of residual-stream tensors accumulate to
~1e-2).  Against the reference-generated fp32 golden (unrounded weights) the bound is 4e-2."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

UNET_EXTRA = dict(  # inference_v2.yaml:1-21
    use_inflated_groupnorm=True, unet_use_cross_frame_attention=False, unet_use_temporal_attention=False,
    use_motion_module=True, motion_module_resolutions=[1, 2, 4, 8], motion_module_mid_block=True,
    motion_module_decoder_only=False, motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=["Temporal_Self", "Temporal_Self"],
                              temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                              temporal_attention_dim_div=1))


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


class _Writer:
    """Stands in for the ReferenceNet write pass: exposes cond banks in pairing order."""
    def __init__(self, banks):
        self.banks = banks


def build_product(cfg, sd, banks_cond, ref_w=0.95, audio_w=3.0):
    from vexpress_b200.modules import ReferenceAttentionControl, UNet3DConditionModel
    model = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"],
                                 cross_attention_dim=cfg["cross_attention_dim"], **UNET_EXTRA)
    model.load_state_dict(sd, strict=True)
    model = model.to(torch.bfloat16).to("cuda")       # cast on the host: the upload is plain memcpys, no cast kernels
    reader = ReferenceAttentionControl(model, do_classifier_free_guidance=True, mode="read", batch_size=1,
                                       fusion_blocks="full", reference_attention_weight=ref_w,
                                       audio_attention_weight=audio_w)
    reader.update(_Writer([b.cuda() for b in banks_cond]), True, dtype=torch.bfloat16)
    return model, reader


def test_unet_small_vs_oracle_and_golden(golden_dir):
    from oracle import vx_oracle as O
    g = torch.load(os.path.join(golden_dir, "unet_small.pt"), weights_only=False)
    cfg = g["cfg"]
    sd = O.synth_state_dict(O.unet_param_shapes(cfg), g["seed_weights"])
    lat, kps, audio, banks = O.synth_inputs(cfg, g["f"], g["h"], g["h"], True, g["seed_inputs"])
    model, reader = build_product(cfg, sd, [b[1:] for b in banks], g["ref_w"], g["audio_w"])
    x = lat.repeat(2, 1, 1, 1, 1)
    enc = audio.reshape(-1, 5, cfg["cross_attention_dim"])

    # oracle on bf16-rounded weights / inputs (isolates activation rounding)
    r = lambda t: t.bfloat16().float()
    sd_r = {k: r(v) for k, v in sd.items()}
    taps_o = {}
    with torch.no_grad():
        ref = O.unet_forward(sd_r, cfg, r(x), 499, r(enc), r(kps), [r(b) for b in banks], g["ref_w"], g["audio_w"], taps=taps_o)

    # product, with taps
    eng = model.engine()
    b, c, f, h, w = x.shape
    frames = x.bfloat16().cuda().permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w).contiguous()
    kps_nhwc = kps.bfloat16().cuda().permute(0, 2, 3, 4, 1).reshape(b * f * h * w, -1).contiguous()
    taps_p = {}
    out_frames = eng.forward_frames(frames, 499, enc.cuda(), kps_nhwc, None, b, f, taps=taps_p)
    torch.cuda.synchronize()
    worst = 0.0
    for k in taps_o:
        if k in taps_p:
            e = _rel(taps_p[k].cpu(), taps_o[k])
            worst = max(worst, e)
            print(f"tap {k:40s} rel={e:.3e}")
    out = out_frames.view(b, f, -1, h, w).permute(0, 2, 1, 3, 4).float().cpu()
    e_or = _rel(out, ref)
    e_gold = _rel(out, g["out_t499"])
    print(f"final: vs oracle(bf16 weights) {e_or:.3e}   vs reference golden(fp32) {e_gold:.3e}   worst tap {worst:.3e}")
    assert worst < 2e-2 and e_or < 3e-2 and e_gold < 4e-2

    # public forward (reference signature) gives the same tensor; second timestep against the golden too
    out2 = model(x.bfloat16().cuda(), 499, encoder_hidden_states=enc.bfloat16().cuda(),
                 kps_features=kps.bfloat16().cuda(), return_dict=False)[0]
    assert out2.shape == (b, 4, f, h, w) and torch.equal(out2.float().cpu(), out)
    out3 = model(x.bfloat16().cuda(), torch.tensor(959), encoder_hidden_states=enc.bfloat16().cuda(),
                 kps_features=kps.bfloat16().cuda()).sample
    e3 = _rel(out3.float().cpu(), g["out_t959"])
    print(f"t=959 vs golden {e3:.3e}")
    assert e3 < 4e-2
    reader.clear()
    with pytest.raises(RuntimeError):
        model(x.bfloat16().cuda(), 499, encoder_hidden_states=enc.bfloat16().cuda(), kps_features=kps.bfloat16().cuda())
