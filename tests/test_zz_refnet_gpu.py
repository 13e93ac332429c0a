"""ReferenceNet write pass on the GPU against the golden produced by the reference's own
UNet2DConditionModel + write hooks (oracle/gen_golden.py:gen_refnet).  Runs after the hot-path suites."""
import os

import pytest
import torch

from oracle import vx_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def test_refnet_write_pass_vs_reference_golden(golden_dir):
    from vexpress_b200.modules import ReferenceAttentionControl, UNet2DConditionModel, UNet3DConditionModel
    from vexpress_b200.modules.unet_3d import attention_block_order
    g = torch.load(os.path.join(golden_dir, "refnet_small.pt"), weights_only=False)
    cfg = g["cfg"]
    net = UNet2DConditionModel(block_out_channels=cfg["block_out_channels"], cross_attention_dim=cfg["cross_attention_dim"])
    net.load_state_dict(O.synth_state_dict(O.refnet_param_shapes(cfg), g["seed_weights"]), strict=True)
    net = net.to(device="cuda", dtype=torch.bfloat16)
    writer = ReferenceAttentionControl(net, do_classifier_free_guidance=True, mode="write", batch_size=1,
                                       fusion_blocks="full")
    x = torch.randn(1, 4, g["h"], g["h"], generator=torch.Generator().manual_seed(g["seed_latents"]))
    enc = torch.zeros(1, 1, cfg["cross_attention_dim"], device="cuda", dtype=torch.bfloat16)
    out = net(x.to(device="cuda", dtype=torch.bfloat16), timestep=0, encoder_hidden_states=enc, return_dict=False)[0]
    torch.cuda.synchronize()
    # hand over to a reader exactly as the pipeline does (v_express_pipeline.py:509)
    unet = UNet3DConditionModel(
        block_out_channels=cfg["block_out_channels"], cross_attention_dim=cfg["cross_attention_dim"],
        use_inflated_groupnorm=True, use_motion_module=True, motion_module_mid_block=True, motion_module_type="Vanilla",
        motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                                  attention_block_types=["Temporal_Self", "Temporal_Self"],
                                  temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                                  temporal_attention_dim_div=1))
    reader = ReferenceAttentionControl(unet, do_classifier_free_guidance=True, mode="read", batch_size=1,
                                       fusion_blocks="full")
    reader.update(writer, True, dtype=torch.bfloat16)
    mods = dict(unet.named_modules())
    names = attention_block_order(unet)
    assert [n.replace(".transformer_blocks.0", "") for n in names] == g["bank_order"]
    worst = 0.0
    for n, ref in zip(names, g["banks"]):
        bank = mods[n].bank
        assert len(bank) == 1 and bank[0].shape == (2,) + tuple(ref.shape[1:]), (n, bank[0].shape)
        assert torch.count_nonzero(bank[0][0]).item() == 0                      # CFG zero half
        err = _rel(bank[0][1:].cpu(), ref)
        worst = max(worst, err)
        assert err < 3e-2, (n, err)                                            # bf16 path vs the fp32 reference
    e_out = _rel(out.cpu(), g["out"])
    print(f"refnet banks worst rel-L2 {worst:.3e}, out rel-L2 {e_out:.3e}")
    assert e_out < 5e-2, e_out
    writer.clear()
    assert all(len(b.bank) == 0 for b in net.writer_blocks())
