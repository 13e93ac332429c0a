"""B200-native ReferenceNet: drop-in for the reference's ``modules/unet_2d_condition.py`` on the one call the
V-Express pipeline makes to it (reference pipelines/v_express_pipeline.py:451-457,501-509):

    writer = ReferenceAttentionControl(reference_net, mode="write", fusion_blocks="full", ...)
    reference_net(ref_latents, timestep=0, encoder_hidden_states=zeros(1, 1, 768), return_dict=False)
    reader.update(writer, do_classifier_free_guidance, dtype=...)

Same constructor configuration (SD-1.5 topology), ``state_dict`` key / shape layout (684 tensors, 859.5 M parameters at
full width -- the reference drops ``conv_norm_out``, modules/unet_2d_condition.py:650) and ``forward`` signature as the
reference ``UNet2DConditionModel`` (modules/unet_2d_condition.py:69-660, 877-1313).  The module tree holds parameters
only; the arithmetic runs on the same sm_100a kernels as the denoising UNet (``UNetEngine``'s resnet / GEMM /
flash-attention / GEGLU schedule) on the channels-last token layout ``[(n)(h w), C]`` with n = 1:

* CrossAttnDownBlock2D / DownBlock2D / UNetMidBlock2DCrossAttn / UpBlock2D / CrossAttnUpBlock2D
  (modules/unet_2d_blocks.py:630-676, 745-775, 470-507, 1027-1073, 890-961);
* Transformer2DModel with conv projections (modules/transformer_2d.py:216-399) around the WRITE branch of the hacked
  ``BasicTransformerBlock.forward`` (modules/mutual_self_attention.py:127-130, 145-174, 270-283): attn1, then
  ``bank.append(norm2(h))``, attn2 against the encoder states, feed-forward.

Module registration order is the reference's (conv_in, time_embedding, down_blocks, up_blocks, mid_block, conv_out):
``ReferenceAttentionControl.update`` pairs writer and reader blocks by a stable sort over depth-first module order.

There is no PyTorch/CPU fallback: ``forward`` raises if the CUDA library is missing.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from .unet_3d import BF16, UNetEngine, _Config, _Node, _Norm

_DOWN2D = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D")
_UP2D = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")


class BasicTransformerBlock(_Node):
    """Parameter holder of one ReferenceNet transformer block + the ``bank`` its write pass fills
    (reference modules/attention.py:12-295; write hook modules/mutual_self_attention.py:165-166)."""

    def __init__(self):
        super().__init__()
        self.bank: List[torch.Tensor] = []


def _unet2d_keys(boc, cross, layers, in_ch, out_ch) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape, inserted in the reference's module registration order."""
    S: Dict[str, Tuple[int, ...]] = {}
    ted = boc[0] * 4

    def wb(p, *shape):
        S[p + ".weight"] = tuple(shape)
        S[p + ".bias"] = (shape[0],)

    def attn(p, c, kv):
        S[p + ".to_q.weight"] = (c, c)
        S[p + ".to_k.weight"] = (c, kv)
        S[p + ".to_v.weight"] = (c, kv)
        wb(p + ".to_out.0", c, c)

    def resnet(p, ci, co):
        wb(p + ".norm1", ci)
        wb(p + ".conv1", co, ci, 3, 3)
        wb(p + ".time_emb_proj", co, ted)
        wb(p + ".norm2", co)
        wb(p + ".conv2", co, co, 3, 3)
        if ci != co:
            wb(p + ".conv_shortcut", co, ci, 1, 1)

    def t2d(p, c):
        wb(p + ".norm", c)
        wb(p + ".proj_in", c, c, 1, 1)
        t = p + ".transformer_blocks.0"
        wb(t + ".norm1", c)
        attn(t + ".attn1", c, c)
        wb(t + ".norm2", c)
        attn(t + ".attn2", c, cross)
        wb(t + ".norm3", c)
        wb(t + ".ff.net.0.proj", 8 * c, c)
        wb(t + ".ff.net.2", c, 4 * c)
        wb(p + ".proj_out", c, c, 1, 1)

    wb("conv_in", boc[0], in_ch, 3, 3)
    wb("time_embedding.linear_1", ted, boc[0])
    wb("time_embedding.linear_2", ted, ted)
    co = boc[0]
    for i in range(4):
        ci, co = co, boc[i]
        for j in range(layers):
            resnet(f"down_blocks.{i}.resnets.{j}", ci if j == 0 else co, co)
            if i < 3:
                t2d(f"down_blocks.{i}.attentions.{j}", co)
        if i < 3:
            wb(f"down_blocks.{i}.downsamplers.0.conv", co, co, 3, 3)
    rev = list(reversed(boc))
    co = rev[0]
    for i in range(4):
        prev, co = co, rev[i]
        ci = rev[min(i + 1, 3)]
        for j in range(layers + 1):
            skip = ci if j == layers else co
            resnet(f"up_blocks.{i}.resnets.{j}", (prev if j == 0 else co) + skip, co)
            if i > 0:
                t2d(f"up_blocks.{i}.attentions.{j}", co)
        if i < 3:
            wb(f"up_blocks.{i}.upsamplers.0.conv", co, co, 3, 3)
    c = boc[-1]
    resnet("mid_block.resnets.0", c, c)
    t2d("mid_block.attentions.0", c)
    resnet("mid_block.resnets.1", c, c)
    wb("conv_out", out_ch, boc[0], 3, 3)
    return S


def _build_tree_2d(root: nn.Module, shapes: Dict[str, Tuple[int, ...]]):
    for key, shape in shapes.items():
        parts = key.split(".")
        node = root
        for depth, name in enumerate(parts[:-1]):
            child = node._modules.get(name)
            if child is None:
                path = parts[:depth + 1]
                if name == "0" and len(path) >= 2 and path[-2] == "transformer_blocks":
                    child = BasicTransformerBlock()
                elif "norm" in name:
                    child = _Norm()
                else:
                    child = _Node()
                node.add_module(name, child)
            node = child
        node.register_parameter(parts[-1], nn.Parameter(torch.empty(shape), requires_grad=False))


def writer_block_names() -> List[str]:
    """The 16 transformer blocks in the reference's depth-first module order (down_blocks, up_blocks, mid_block)."""
    dfs = [f"down_blocks.{i}.attentions.{j}" for i in range(3) for j in range(2)]
    dfs += [f"up_blocks.{i}.attentions.{j}" for i in (1, 2, 3) for j in range(3)]
    dfs += ["mid_block.attentions.0"]
    return [n + ".transformer_blocks.0" for n in dfs]


class UNet2DConditionModel(nn.Module):
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0, down_block_types=_DOWN2D, mid_block_type="UNetMidBlock2DCrossAttn",
                 up_block_types=_UP2D, only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280),
                 layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1, act_fn="silu",
                 norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280, attention_head_dim=8,
                 dual_cross_attention=False, use_linear_projection=False, class_embed_type=None,
                 num_class_embeds=None, upcast_attention=False, resnet_time_scale_shift="default", **unused):
        super().__init__()
        cfg = dict(locals())
        for k in ("self", "__class__", "unused", "cfg"):
            cfg.pop(k, None)
        self.config = _Config(cfg)
        self.sample_size = sample_size

        def need(cond, what):
            if not cond:
                raise ValueError(f"vexpress_b200.UNet2DConditionModel: unsupported configuration: {what}")
        need(mid_block_type == "UNetMidBlock2DCrossAttn", "mid block type")
        need(tuple(down_block_types) == _DOWN2D and tuple(up_block_types) == _UP2D, "block types")
        need(len(block_out_channels) == 4 and layers_per_block == 2, "4 levels x 2 layers")
        need(act_fn in ("silu", "swish") and norm_num_groups == 32 and resnet_time_scale_shift == "default", "act/norm")
        need(not center_input_sample and flip_sin_to_cos and freq_shift == 0, "time projection")
        need(not dual_cross_attention and not use_linear_projection and class_embed_type is None
             and num_class_embeds is None and not only_cross_attention and not upcast_attention, "attention flags")
        need(downsample_padding == 1 and mid_block_scale_factor == 1, "downsample padding / output scale")
        need(attention_head_dim == 8 or tuple(attention_head_dim) == (8, 8, 8, 8), "attention_head_dim (= #heads) 8")
        need(all(c % 64 == 0 for c in block_out_channels), "channel widths must be multiples of 64")
        self.heads = 8
        _build_tree_2d(self, _unet2d_keys(tuple(block_out_channels), cross_attention_dim, layers_per_block, in_channels,
                                          out_channels))
        self._engine: Optional[RefNetEngine] = None
        self.write_banks = False      # set by ReferenceAttentionControl(mode="write")

    @property
    def in_channels(self):
        return self.config["in_channels"]

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    @classmethod
    def load_config(cls, path):
        with open(path) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, **kwargs):
        if not isinstance(config, dict):
            config = cls.load_config(config)
        init = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        init.update(kwargs)
        return cls(**init)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def writer_blocks(self) -> List[BasicTransformerBlock]:
        mods = dict(self.named_modules())
        return [mods[n] for n in writer_block_names()]

    def engine(self) -> "RefNetEngine":
        if self._engine is None:
            self._engine = RefNetEngine(self)
        return self._engine

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, added_cond_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None,
                encoder_attention_mask=None, return_dict: bool = True):
        """Reference signature (modules/unet_2d_condition.py:877-892).  sample (n,4,h,w); encoder_hidden_states
        (n, L<=8, cross).  Returns the sample (n,4,h,w) in the model dtype; with a write-mode control installed every
        transformer block appends its ``norm2`` activations (n, h*w, C) to ``block.bank``."""
        if any(v is not None for v in (class_labels, timestep_cond, attention_mask, cross_attention_kwargs,
                                       added_cond_kwargs, down_block_additional_residuals,
                                       mid_block_additional_residual, encoder_attention_mask)):
            raise ValueError("only (sample, timestep, encoder_hidden_states) are used on the V-Express path")
        assert sample.dim() == 4, f"Expected sample to have ndim=4, but got ndim={sample.dim()}."
        out = self.engine().forward(sample.to(BF16).contiguous(), timestep, encoder_hidden_states).to(sample.dtype)
        if not return_dict:
            return (out,)
        from .unet_3d import UNet3DConditionOutput
        return UNet3DConditionOutput(sample=out)


class RefNetEngine(UNetEngine):
    """Packed weights + kernel schedule of the ReferenceNet; block kernels are inherited from ``UNetEngine``."""

    def __init__(self, model: UNet2DConditionModel):  # noqa: D107 (does not call UNetEngine.__init__: no banks to read)
        from .. import _ffi
        _ffi.require_sm100()
        dev = model.device
        if dev.type != "cuda":
            raise RuntimeError("vexpress_b200: the model must live on a CUDA (sm_100a) device; there is no CPU path")
        _ffi.note_compute_dtype(model.dtype, "UNet2DConditionModel")
        self.model = model
        self.dev = dev
        cfg = model.config
        self.boc = tuple(cfg["block_out_channels"])
        self.heads = model.heads
        self.groups = cfg["norm_num_groups"]
        self.eps = float(cfg["norm_eps"])
        self.cross = cfg["cross_attention_dim"]
        self.sd = {k: v.detach() for k, v in model.state_dict().items()}
        self.W: Dict[str, torch.Tensor] = {}
        self._pack(self.sd)
        self.ln_fold = self.ln_fuse = False   # the write pass materialises LayerNorm(h) (it IS the bank): plain LayerNorm kernels

    def _transformer_write(self, p, x, NB, HW, enc_flat):
        """GroupNorm -> proj_in -> [norm1, attn1] -> bank = norm2(h) -> attn2(enc) -> ff -> proj_out + residual."""
        W = self.W
        C = x.shape[1]
        heads = self.heads
        h = ops.groupnorm(x, NB, HW, W[p + ".norm.weight"], W[p + ".norm.bias"], 1e-6, False, groups=self.groups)
        h = ops.gemm(h, W[p + ".proj_in.weight"], W[p + ".proj_in.bias"])
        t = p + ".transformer_blocks.0"
        n = ops.layernorm(h, W[t + ".norm1.weight"], W[t + ".norm1.bias"])
        qkv = ops.gemm(n, W[t + ".attn1.qkv"])
        a = ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], heads, HW, HW)
        h = ops.gemm(a, W[t + ".attn1.to_out.0.weight"], W[t + ".attn1.to_out.0.bias"], residual=h)
        n = ops.layernorm(h, W[t + ".norm2.weight"], W[t + ".norm2.bias"])
        if self.model.write_banks:
            self.model.get_submodule(t).bank.append(n.view(NB, HW, C).clone())
        q = ops.gemm(n, W[t + ".attn2.to_q.weight"])
        kv = ops.gemm(enc_flat, W[t + ".attn2.kv"])
        Lk = enc_flat.shape[0] // NB
        a = ops.smallkv_attention(q, kv[:, :C], kv[:, C:], HW, heads, Lk)
        h = ops.gemm(a, W[t + ".attn2.to_out.0.weight"], W[t + ".attn2.to_out.0.bias"], residual=h)
        n = ops.layernorm(h, W[t + ".norm3.weight"], W[t + ".norm3.bias"])
        h = self._ff(t + ".ff", n, h)
        return ops.gemm(h, W[p + ".proj_out.weight"], W[p + ".proj_out.bias"], residual=x)

    def forward(self, frames, timestep, enc):
        """frames (n,4,h,w) bf16; enc (n, L, cross).  Returns (n,4,h,w) bf16."""
        W = self.W
        NB, cin, H, Wd = frames.shape
        if enc.shape[0] != NB or enc.shape[1] > 8:
            raise ValueError(f"encoder_hidden_states {tuple(enc.shape)}: need batch {NB} and at most 8 tokens")
        boc = self.boc
        temb = self.time_embedding(timestep)
        enc_flat = enc.to(device=self.dev, dtype=BF16).reshape(-1, enc.shape[-1]).contiguous()
        x = ops.conv_in(frames, W["conv_in.weight"], W["conv_in.bias"], boc[0])
        skips = [(x, H, Wd)]
        h_, w_ = H, Wd
        for i in range(4):
            p = f"down_blocks.{i}"
            for j in range(2):
                x = self._resnet(f"{p}.resnets.{j}", x, None, NB, h_, w_, temb)
                if i < 3:
                    x = self._transformer_write(f"{p}.attentions.{j}", x, NB, h_ * w_, enc_flat)
                skips.append((x, h_, w_))
            if i < 3:
                x = ops.downsample_conv(x, NB, h_, w_, W[f"{p}.downsamplers.0.conv.weight"], W[f"{p}.downsamplers.0.conv.bias"])
                h_, w_ = h_ // 2, w_ // 2
                skips.append((x, h_, w_))
        x = self._resnet("mid_block.resnets.0", x, None, NB, h_, w_, temb)
        x = self._transformer_write("mid_block.attentions.0", x, NB, h_ * w_, enc_flat)
        x = self._resnet("mid_block.resnets.1", x, None, NB, h_, w_, temb)
        for i in range(4):
            p = f"up_blocks.{i}"
            for j in range(3):
                skip, sh, sw = skips.pop()
                assert (sh, sw) == (h_, w_)
                x = self._resnet(f"{p}.resnets.{j}", x, skip, NB, h_, w_, temb)
                if i > 0:
                    x = self._transformer_write(f"{p}.attentions.{j}", x, NB, h_ * w_, enc_flat)
            if i < 3:
                x = ops.upconv3x3(x.view(NB, h_, w_, -1), W[f"{p}.upsamplers.0.conv.weight"], W[f"{p}.upsamplers.0.conv.bias"])
                h_, w_ = 2 * h_, 2 * w_
        # no conv_norm_out / activation: the reference resets conv_norm_out to None (unet_2d_condition.py:650,1301-1304)
        out = torch.empty((NB, self.model.config["out_channels"], H, Wd), device=self.dev, dtype=BF16)
        ops.conv_out_tc(x, NB, H, Wd, W["conv_out.packed_w"], W["conv_out.packed_b"], out)
        return out
