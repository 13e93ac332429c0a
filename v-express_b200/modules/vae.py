"""B200-native VAE decoder: drop-in for the ``AutoencoderKL.decode`` the reference pipeline calls
(pipelines/v_express_pipeline.py:152-166; diffusers 0.29.2 AutoencoderKL with the sd-vae-ft-mse config,
SURVEY.md Appendix B.6).  Same ``state_dict`` keys as diffusers for ``post_quant_conv.*`` and ``decoder.*``
(encoder keys are accepted and ignored: VAE *encode* is outside the hot path).

All frames are decoded in ONE batch on the channels-last bf16 layout (the reference decodes frame by frame with
a device->host copy per frame): two-source-free GroupNorm+SiLU kernels, tcgen05 implicit-GEMM 3x3 convs with the
residual add in the epilogue, nearest-2x upsample, and the single-head hd=512 mid-block attention as
GEMM(QK^T, fp32 scores) -> row softmax -> GEMM(PV).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import nn

from .. import ops
from .unet_3d import _Config, _Node, _Norm

BF16 = torch.bfloat16


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


def _vae_keys(boc, layers, latent, out_ch) -> Dict[str, Tuple[int, ...]]:
    S: Dict[str, Tuple[int, ...]] = {}

    def wb(p, *shape):
        S[p + ".weight"] = tuple(shape)
        S[p + ".bias"] = (shape[0],)

    def res(p, ci, co):
        wb(p + ".norm1", ci)
        wb(p + ".conv1", co, ci, 3, 3)
        wb(p + ".norm2", co)
        wb(p + ".conv2", co, co, 3, 3)
        if ci != co:
            wb(p + ".conv_shortcut", co, ci, 1, 1)

    wb("post_quant_conv", latent, latent, 1, 1)
    d = "decoder"
    top = boc[-1]
    wb(d + ".conv_in", top, latent, 3, 3)
    res(d + ".mid_block.resnets.0", top, top)
    a = d + ".mid_block.attentions.0"
    wb(a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        wb(f"{a}.{n}", top, top)
    res(d + ".mid_block.resnets.1", top, top)
    rev = list(reversed(boc))
    co = rev[0]
    for i in range(len(boc)):
        ci, co = co, rev[i]
        for j in range(layers + 1):
            res(f"{d}.up_blocks.{i}.resnets.{j}", ci if j == 0 else co, co)
        if i < len(boc) - 1:
            wb(f"{d}.up_blocks.{i}.upsamplers.0.conv", co, co, 3, 3)
    wb(d + ".conv_norm_out", boc[0])
    wb(d + ".conv_out", out_ch, boc[0], 3, 3)
    return S


class AutoencoderKL(nn.Module):
    """Decoder half of diffusers' AutoencoderKL (parameters only; arithmetic on sm_100a kernels)."""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, act_fn="silu", scaling_factor=0.18215, **_ignored):
        super().__init__()
        if norm_num_groups != 32 or act_fn != "silu" or any(c % 64 for c in block_out_channels):
            raise ValueError("vexpress_b200.AutoencoderKL: unsupported configuration")
        self.config = _Config(in_channels=in_channels, out_channels=out_channels,
                              block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                              latent_channels=latent_channels, norm_num_groups=norm_num_groups,
                              scaling_factor=scaling_factor)
        for key, shape in _vae_keys(tuple(block_out_channels), layers_per_block, latent_channels, out_channels).items():
            parts = key.split(".")
            node = self
            for depth, name in enumerate(parts[:-1]):
                child = node._modules.get(name)
                if child is None:
                    child = _Norm() if "norm" in name else _Node()
                    node.add_module(name, child)
                node = child
            node.register_parameter(parts[-1], nn.Parameter(torch.empty(shape), requires_grad=False))
        self._engine: Optional[VaeDecoderEngine] = None

    @property
    def dtype(self):
        return self.post_quant_conv.weight.dtype

    @property
    def device(self):
        return self.post_quant_conv.weight.device

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._engine = None
        sd = {}
        # checkpoints saved before diffusers 0.18 (stabilityai/sd-vae-ft-mse among them) name the mid-block attention
        # projections query / key / value / proj_attn; diffusers renames them at load time
        # (AutoencoderKL._convert_deprecated_attention_blocks) -- same here
        old = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}
        for k, v in state_dict.items():
            if k.startswith("encoder.") or k.startswith("quant_conv."):
                continue
            if ".attentions." in k:
                for a, b in old.items():
                    if a in k:
                        k = k.replace(a, b)
                        if v.dim() == 4:
                            v = v.reshape(v.shape[0], v.shape[1])
            sd[k] = v
        return super().load_state_dict(sd, strict=strict, **kw)

    def engine(self) -> "VaeDecoderEngine":
        if self._engine is None:
            self._engine = VaeDecoderEngine(self)
        return self._engine

    def encode(self, *a, **k):
        raise NotImplementedError("VAE encode is outside the denoising hot path (SURVEY.md 8f-f4)")

    @torch.no_grad()
    def decode(self, z, return_dict: bool = True):
        """z (n,4,h,w), already divided by the scaling factor by the caller -> sample (n,3,8h,8w), model dtype."""
        out = self.engine().decode(z.to(BF16), pre_scale=1.0, post=False, out_dtype=BF16).to(z.dtype)
        return DecoderOutput(out) if return_dict else (out,)

    @torch.no_grad()
    def decode_latents(self, latents, out=None):
        """Fused form of the reference's ``decode_latents`` (pipelines/v_express_pipeline.py:152-166):
        latents (n,4,h,w) UNSCALED -> fp32 frames (n,3,8h,8w) in [0,1] on the device.  ``out``: optional fp32 view of
        that shape whose (h, w) planes are contiguous (frame / channel strides are free)."""
        return self.engine().decode(latents.to(BF16), pre_scale=1.0 / self.config["scaling_factor"], post=True,
                                    out_dtype=torch.float32, out=out)


class VaeDecoderEngine:
    def __init__(self, model: AutoencoderKL):
        from .. import _ffi
        _ffi.require_sm100()
        if model.device.type != "cuda":
            raise RuntimeError("vexpress_b200: the VAE must live on a CUDA (sm_100a) device; there is no CPU path")
        _ffi.note_compute_dtype(model.dtype, "AutoencoderKL")
        self.model = model
        self.dev = model.device
        self.boc = tuple(model.config["block_out_channels"])
        self.layers = model.config["layers_per_block"]
        self.W: Dict[str, torch.Tensor] = {}
        dev = self.dev
        msd = model.state_dict()
        self.W.update(ops.f32_arena({k: v for k, v in msd.items() if k.endswith(".bias") or v.dim() == 1}, dev))
        for k, v in msd.items():
            v = v.detach()
            p = k.rsplit(".", 1)[0]
            if k.endswith(".bias") or v.dim() == 1:
                continue
            elif p == "decoder.conv_in":
                self.W[k] = v.to(device=dev, dtype=BF16).float().reshape(v.shape[0], -1).t().contiguous()
            elif p == "decoder.conv_out":
                self.W["conv_out.packed_w"], self.W["conv_out.packed_b"] = ops.pack_conv_out(
                    v.to(device=dev, dtype=BF16), self.W[p + ".bias"])
            elif p == "post_quant_conv":
                self.W[k] = v.to(device=dev, dtype=BF16).float().reshape(v.shape[0], v.shape[1]).contiguous()
            elif v.dim() == 4 and v.shape[-1] == 3 and ".upsamplers." in k:
                self.W[k] = ops.pack_upconv_weight(v.to(device=dev, dtype=BF16))    # upsample folded into the conv
            elif v.dim() == 4 and v.shape[-1] == 3:
                self.W[k] = ops.pack_conv3x3_weight(v.to(device=dev, dtype=BF16))
            elif v.dim() == 4:
                self.W[k] = v.to(device=dev, dtype=BF16).reshape(v.shape[0], v.shape[1]).contiguous()
            else:
                self.W[k] = v.to(device=dev, dtype=BF16).contiguous()

    def _res(self, p, x, NB, H, Wd):
        W = self.W
        h = ops.groupnorm(x, NB, H * Wd, W[p + ".norm1.weight"], W[p + ".norm1.bias"], 1e-6, True)
        h = ops.conv3x3(h.view(NB, H, Wd, -1), W[p + ".conv1.weight"], W[p + ".conv1.bias"])
        h = ops.groupnorm(h, NB, H * Wd, W[p + ".norm2.weight"], W[p + ".norm2.bias"], 1e-6, True)
        sc = x
        if (p + ".conv_shortcut.weight") in W:
            sc = ops.gemm(x, W[p + ".conv_shortcut.weight"], W[p + ".conv_shortcut.bias"])
        return ops.conv3x3(h.view(NB, H, Wd, -1), W[p + ".conv2.weight"], W[p + ".conv2.bias"], residual=sc)

    def _attn(self, p, x, NB, HW):
        """Single-head attention over the HW tokens of each frame (diffusers Attention with group_norm,
        biased projections, residual connection)."""
        W = self.W
        C = x.shape[1]
        xn = ops.groupnorm(x, NB, HW, W[p + ".group_norm.weight"], W[p + ".group_norm.bias"], 1e-6, False)
        q = ops.gemm(xn, W[p + ".to_q.weight"], W[p + ".to_q.bias"])
        k = ops.gemm(xn, W[p + ".to_k.weight"], W[p + ".to_k.bias"])
        o = torch.empty_like(q)
        scores = torch.empty((HW, HW), device=self.dev, dtype=torch.float32)
        probs = torch.empty((HW, HW), device=self.dev, dtype=BF16)
        vt = torch.empty((C, HW), device=self.dev, dtype=BF16)
        for n in range(NB):
            sl = slice(n * HW, (n + 1) * HW)
            ops.gemm(q[sl], k[sl], scale=C ** -0.5, out=scores, out_f32=True)
            ops.softmax_rows(scores, out=probs)
            # V^T = Wv @ xn^T (bias added after P@V: softmax rows sum to 1, so P (V + 1 b^T) = P V + b^T)
            ops.gemm(W[p + ".to_v.weight"], xn[sl], out=vt)
            ops.gemm(probs, vt, W[p + ".to_v.bias"], out=o[sl])
        return ops.gemm(o, W[p + ".to_out.0.weight"], W[p + ".to_out.0.bias"], residual=x)

    def decode(self, z, pre_scale: float, post: bool, out_dtype, out=None):
        W = self.W
        assert z.dim() == 4
        z = z.contiguous()
        NB, _, H, Wd = z.shape
        d = "decoder"
        x = ops.conv_in(z, W[d + ".conv_in.weight"], W[d + ".conv_in.bias"], self.boc[-1], pre_scale=pre_scale,
                        pre_w=W["post_quant_conv.weight"], pre_b=W["post_quant_conv.bias"])
        x = self._res(d + ".mid_block.resnets.0", x, NB, H, Wd)
        x = self._attn(d + ".mid_block.attentions.0", x, NB, H * Wd)
        x = self._res(d + ".mid_block.resnets.1", x, NB, H, Wd)
        nb = len(self.boc)
        for i in range(nb):
            for j in range(self.layers + 1):
                x = self._res(f"{d}.up_blocks.{i}.resnets.{j}", x, NB, H, Wd)
            if i < nb - 1:
                x = ops.upconv3x3(x.view(NB, H, Wd, -1), W[f"{d}.up_blocks.{i}.upsamplers.0.conv.weight"],
                                  W[f"{d}.up_blocks.{i}.upsamplers.0.conv.bias"])
                H, Wd = 2 * H, 2 * Wd
        x = ops.groupnorm(x, NB, H * Wd, W[d + ".conv_norm_out.weight"], W[d + ".conv_norm_out.bias"], 1e-6, True)
        if out is None:
            out = torch.empty((NB, self.model.config["out_channels"], H, Wd), device=self.dev, dtype=out_dtype)
        assert out.shape == (NB, self.model.config["out_channels"], H, Wd) and out.dtype == out_dtype
        ops.conv_out_tc(x, NB, H, Wd, W["conv_out.packed_w"], W["conv_out.packed_b"], out, post=post)
        return out
