"""CPU oracle for the V-Express denoising hot path (TEST INFRASTRUCTURE ONLY).

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``vexpress_b200``) never imports anything under ``oracle/`` and fails loudly when the CUDA
library is missing.

It is a plain fp32 PyTorch restatement of the reference algorithm, written functionally over a
``state_dict`` in the reference's key layout.  Each function cites the
reference file:line it follows (paths relative to the upstream repo).  Third-party leaves that are
not vendored in the reference (``diffusers==0.29.2``: Attention/AttnProcessor2_0, FeedForward/GEGLU,
Timesteps/TimestepEmbedding, DDIMScheduler, AutoencoderKL decoder) are restated from their published
algorithm.

Parity pinning: the reference ships no tests or golden vectors for this path ("parity unpinned" by
the reference itself).  The pins are therefore generated *from the reference's own code run
verbatim* in the build container (``oracle/gen_golden.py`` imports ``/root/reference/modules/*.py``
and ``pipelines/*.py`` over the small ``oracle/diffusers_shim``) and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# Topology (reference: modules/unet_3d.py:34-250 with inference_v2.yaml:1-21 and the SD-1.5
# unet/config.json)
# --------------------------------------------------------------------------------------

DEFAULT_CFG = dict(
    in_channels=4,
    out_channels=4,
    block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2,
    heads=8,                    # "attention_head_dim: 8" is used as the number of heads (unet_3d_blocks.py:353-356)
    cross_attention_dim=768,
    norm_num_groups=32,
    norm_eps=1e-5,
    temporal_max_len=32,
)


def small_cfg(width=(64, 128, 256, 256), cross=768):
    c = dict(DEFAULT_CFG)
    c["block_out_channels"] = tuple(width)
    c["cross_attention_dim"] = cross
    return c


def attention_block_names() -> List[str]:
    """Names of the 16 spatial transformer blocks in ``torch_dfs(unet)`` order.

    Module registration order of UNet3DConditionModel is conv_in, time_proj, time_embedding,
    down_blocks, up_blocks, mid_block, ... (``self.mid_block = None`` at unet_3d.py:110 is a plain
    attribute; the module is registered only at :160, after ``up_blocks``)."""
    names = []
    for i in range(3):
        for j in range(2):
            names.append(f"down_blocks.{i}.attentions.{j}")
    for i in (1, 2, 3):
        for j in range(3):
            names.append(f"up_blocks.{i}.attentions.{j}")
    names.append("mid_block.attentions.0")
    return names


def attention_block_dim(name: str, cfg) -> int:
    boc = cfg["block_out_channels"]
    if name.startswith("down_blocks."):
        return boc[int(name.split(".")[1])]
    if name.startswith("up_blocks."):
        return list(reversed(boc))[int(name.split(".")[1])]
    return boc[-1]


def bank_order(cfg=DEFAULT_CFG) -> List[str]:
    """Pairing order of reader blocks: stable sort of dfs order by -norm1 dim
    (modules/mutual_self_attention.py:346-351)."""
    names = attention_block_names()
    return sorted(names, key=lambda n: -attention_block_dim(n, cfg))


# --------------------------------------------------------------------------------------
# Context scheduler (pipelines/context.py:22-66) -- integer, bit-exact
# --------------------------------------------------------------------------------------

def ordered_halving(val: int) -> float:
    """pipelines/context.py:22-27: bit-reverse a 64-bit integer and scale to [0,1)."""
    rev = int(f"{val:064b}"[::-1], 2)
    return rev / (1 << 64)


def uniform(step, num_frames, context_size, context_stride=3, context_overlap=4, closed_loop=True):
    """pipelines/context.py:30-59."""
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    context_stride = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    for context_step in 1 << np.arange(context_stride):
        pad = int(round(num_frames * ordered_halving(step)))
        for j in range(
            int(ordered_halving(step) * context_step) + pad,
            num_frames + pad + (0 if closed_loop else -context_overlap),
            (context_size * context_step - context_overlap),
        ):
            window = []
            for e in range(j, j + context_size * context_step, context_step):
                if e >= num_frames:
                    e = num_frames - 2 - e % num_frames
                window.append(int(e))
            yield window


def get_context_scheduler(name: str):
    """pipelines/context.py:62-66."""
    if name == "uniform":
        return uniform
    raise ValueError(f"Unknown context_overlap policy {name}")


def context_windows(video_length, context_frames, context_overlap):
    """The call the pipeline makes (pipelines/v_express_pipeline.py:486-496)."""
    return list(uniform(0, video_length, context_frames, 1, context_overlap, False))


def num_frame_context(windows, video_length) -> np.ndarray:
    """pipelines/v_express_pipeline.py:498-500: index-put ``+= 1`` is NON-accumulating for duplicate
    indices inside one window (a duplicated frame counts once per window)."""
    cnt = np.zeros(video_length, dtype=np.int64)
    for w in windows:
        cnt[np.unique(np.asarray(w, dtype=np.int64))] += 1
    return cnt


# --------------------------------------------------------------------------------------
# DDIM scheduler (diffusers 0.29.2 DDIMScheduler with inference_v2.yaml:23-33)
# --------------------------------------------------------------------------------------

class DDIM:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=True,
                 steps_offset=1, prediction_type="v_prediction", rescale_betas_zero_snr=True,
                 timestep_spacing="trailing", **_):
        assert beta_schedule == "scaled_linear" and prediction_type == "v_prediction"
        assert timestep_spacing == "trailing" and not clip_sample
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        if rescale_betas_zero_snr:
            alphas = 1.0 - betas
            abar_sqrt = torch.cumprod(alphas, dim=0).sqrt()
            a0 = abar_sqrt[0].clone()
            aT = abar_sqrt[-1].clone()
            abar_sqrt -= aT
            abar_sqrt *= a0 / (a0 - aT)
            abar = abar_sqrt ** 2
            alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
            betas = 1 - alphas
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_train_timesteps = num_train_timesteps
        self.order = 1
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps / num_inference_steps
        ts = np.round(np.arange(self.num_train_timesteps, 0, -step_ratio)).astype(np.int64) - 1
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coeffs(self, timestep: int):
        """fp32 scalars (a_t, a_prev) used by step()."""
        prev = int(timestep) - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[int(timestep)]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return a_t, a_prev

    def step(self, model_output, timestep, sample, eta=0.0, **_):
        assert eta == 0.0
        a_t, a_prev = self.coeffs(int(timestep))
        b_t = 1 - a_t
        x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
        eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        direction = (1 - a_prev) ** 0.5 * eps
        prev_sample = a_prev ** 0.5 * x0 + direction
        return _StepOut(prev_sample)


class _StepOut:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


# --------------------------------------------------------------------------------------
# diffusers leaves
# --------------------------------------------------------------------------------------

def timestep_embedding(t: Tensor, dim=320) -> Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): fp32 [cos | sin]."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def attention(sd, p, x, ctx, heads):
    """diffusers Attention + AttnProcessor2_0 (Appendix B.1/B.2): no-bias q/k/v, SDPA, biased out."""
    q = _lin(sd, p + ".to_q", x)
    k = _lin(sd, p + ".to_k", ctx)
    v = _lin(sd, p + ".to_v", ctx)
    B, Lq, inner = q.shape
    hd = inner // heads
    q = q.view(B, Lq, heads, hd).transpose(1, 2)
    k = k.view(B, -1, heads, hd).transpose(1, 2)
    v = v.view(B, -1, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, Lq, inner)
    return _lin(sd, p + ".to_out.0", o)


def feed_forward(sd, p, x):
    """diffusers FeedForward(geglu) (Appendix B.3): proj -> (h, gate) -> h*gelu_erf(gate) -> linear."""
    h = _lin(sd, p + ".net.0.proj", x)
    h, gate = h.chunk(2, dim=-1)
    return _lin(sd, p + ".net.2", h * F.gelu(gate))


def layer_norm(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def group_norm(sd, p, x, groups, eps):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


# --------------------------------------------------------------------------------------
# UNet3D blocks on ((b f), c, h, w) frames
# --------------------------------------------------------------------------------------

def resnet_block(sd, p, x, emb_bf, groups, eps):
    """modules/resnet.py:217-251 (ResnetBlock3D.forward; per-frame GN = InflatedGroupNorm :20-28).
    ``emb_bf``: (B, 1280) time embedding already broadcast to frames."""
    h = group_norm(sd, p + ".norm1", x, groups, eps)
    h = F.silu(h)
    h = conv(sd, p + ".conv1", h)
    if emb_bf is not None:
        t = _lin(sd, p + ".time_emb_proj", F.silu(emb_bf))
        h = h + t[:, :, None, None]
    h = group_norm(sd, p + ".norm2", h, groups, eps)
    h = F.silu(h)
    h = conv(sd, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in sd:
        x = conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def spatial_transformer(sd, p, x, enc, bank, heads, groups, ref_w, audio_w, f):
    """modules/transformer_3d.py:103-169 + read branch of the hacked block forward
    (modules/mutual_self_attention.py:101-131,176-267).

    x (B,C,h,w); enc (B,5,768); bank (b,N,C) -> repeated over f frames (:205-213)."""
    B, C, H, W = x.shape
    res = x
    h = group_norm(sd, p + ".norm", x, groups, 1e-6)
    h = conv(sd, p + ".proj_in", h, padding=0)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    tb = p + ".transformer_blocks.0"
    n = layer_norm(sd, tb + ".norm1", h)
    h = attention(sd, tb + ".attn1", n, n, heads) + h
    n = layer_norm(sd, tb + ".norm1_5", h)
    bank_bf = bank.unsqueeze(1).repeat(1, f, 1, 1).reshape(B, bank.shape[1], bank.shape[2])
    a = attention(sd, tb + ".attn1_5", n, bank_bf, heads)
    if ref_w != 1.0:
        a = a * ref_w
    h = a + h
    n = layer_norm(sd, tb + ".norm2", h)
    a = attention(sd, tb + ".attn2", n, enc, heads)
    if audio_w != 1.0:
        a = a * audio_w
    h = a + h
    h = feed_forward(sd, tb + ".ff", layer_norm(sd, tb + ".norm3", h)) + h
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    h = conv(sd, p + ".proj_out", h, padding=0)
    return h + res


def positional_encoding(d_model, max_len) -> Tensor:
    """modules/motion_module.py:262-273."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def motion_module(sd, p, x, heads, groups, f):
    """modules/motion_module.py:146-182 (TemporalTransformer3DModel.forward), :236-259 (block),
    :351-388 (VersatileAttention: ((b f) d c)->((b d) f c), +PE, self-attention over f)."""
    p = p + ".temporal_transformer"
    B, C, H, W = x.shape
    b = B // f
    d = H * W
    res = x
    h = group_norm(sd, p + ".norm", x, groups, 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(B, d, C)
    h = _lin(sd, p + ".proj_in", h)
    tb = p + ".transformer_blocks.0"
    for i in range(2):
        n = layer_norm(sd, tb + f".norms.{i}", h)
        n = n.view(b, f, d, C).permute(0, 2, 1, 3).reshape(b * d, f, C)
        pe = sd[tb + f".attention_blocks.{i}.pos_encoder.pe"]
        n = n + pe[:, :f]
        a = attention(sd, tb + f".attention_blocks.{i}", n, n, heads)
        a = a.view(b, d, f, C).permute(0, 2, 1, 3).reshape(B, d, C)
        h = a + h
    h = feed_forward(sd, tb + ".ff", layer_norm(sd, tb + ".ff_norm", h)) + h
    h = _lin(sd, p + ".proj_out", h)
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return h + res


def unet_forward(sd: Dict[str, Tensor], cfg, sample: Tensor, timestep, enc: Tensor,
                 kps_features: Optional[Tensor], banks: Sequence[Tensor],
                 ref_w: float = 1.0, audio_w: float = 1.0, taps: Optional[dict] = None) -> Tensor:
    """modules/unet_3d.py:400-578 (UNet3DConditionModel.forward) with the read-mode hooks installed
    and ``banks`` given in pairing order (``bank_order``).  sample (b,4,f,h,w) -> (b,4,f,h,w).

    ``taps`` (optional dict) receives named intermediate activations as ((b f),c,h,w) tensors."""
    b, cin, f, H, W = sample.shape
    B = b * f
    boc = cfg["block_out_channels"]
    heads, groups, eps = cfg["heads"], cfg["norm_num_groups"], cfg["norm_eps"]
    order = bank_order(cfg)
    bank_of = {name: banks[i] for i, name in enumerate(order)}

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    # time (unet_3d.py:449-470): timesteps.expand(b); fp32 sinusoid; cast; MLP
    t = torch.as_tensor(timestep).reshape(-1)[:1].expand(b)
    t_emb = timestep_embedding(t, boc[0]).to(sample.dtype)
    emb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", t_emb)))
    emb_bf = emb.repeat_interleave(f, dim=0)  # temb has batch b, broadcast over f (resnet.py:227-228)

    x = sample.permute(0, 2, 1, 3, 4).reshape(B, cin, H, W)
    x = conv(sd, "conv_in", x)
    if kps_features is not None:
        x = x + kps_features.permute(0, 2, 1, 3, 4).reshape(B, boc[0], H, W)
    tap("conv_in", x)
    skips = [x]

    def attn(p, x):
        return spatial_transformer(sd, p, x, enc, bank_of[p], heads, groups, ref_w, audio_w, f)

    # down (unet_3d_blocks.py:398-461, 537-580)
    for i in range(4):
        p = f"down_blocks.{i}"
        for j in range(cfg["layers_per_block"]):
            x = resnet_block(sd, f"{p}.resnets.{j}", x, emb_bf, groups, eps)
            tap(f"{p}.resnets.{j}", x)
            if i < 3:
                x = attn(f"{p}.attentions.{j}", x)
                tap(f"{p}.attentions.{j}", x)
            x = motion_module(sd, f"{p}.motion_modules.{j}", x, heads, groups, f)
            tap(f"{p}.motion_modules.{j}", x)
            skips.append(x)
        if i < 3:
            x = conv(sd, f"{p}.downsamplers.0.conv", x, stride=2, padding=1)
            tap(f"{p}.downsamplers.0", x)
            skips.append(x)

    # mid (unet_3d_blocks.py:269-293)
    x = resnet_block(sd, "mid_block.resnets.0", x, emb_bf, groups, eps)
    x = attn("mid_block.attentions.0", x)
    x = motion_module(sd, "mid_block.motion_modules.0", x, heads, groups, f)
    x = resnet_block(sd, "mid_block.resnets.1", x, emb_bf, groups, eps)
    tap("mid_block", x)

    # up (unet_3d_blocks.py:679-749, 819-866): cat(skip) -> resnet -> (attn) -> motion; upsample
    for i in range(4):
        p = f"up_blocks.{i}"
        for j in range(cfg["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"{p}.resnets.{j}", x, emb_bf, groups, eps)
            tap(f"{p}.resnets.{j}", x)
            if i > 0:
                x = attn(f"{p}.attentions.{j}", x)
                tap(f"{p}.attentions.{j}", x)
            x = motion_module(sd, f"{p}.motion_modules.{j}", x, heads, groups, f)
            tap(f"{p}.motion_modules.{j}", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")   # resnet.py:53-82
            x = conv(sd, f"{p}.upsamplers.0.conv", x)
            tap(f"{p}.upsamplers.0", x)

    # out (unet_3d.py:571-573)
    x = group_norm(sd, "conv_norm_out", x, groups, eps)
    x = F.silu(x)
    x = conv(sd, "conv_out", x)
    return x.view(b, f, -1, H, W).permute(0, 2, 1, 3, 4).contiguous()


# --------------------------------------------------------------------------------------
# ReferenceNet write pass: SD-1.5 UNet2DConditionModel at t = 0 with a zero text token,
# every BasicTransformerBlock appending norm2(h) to its bank
# --------------------------------------------------------------------------------------

def transformer_2d_write(sd, p, x, enc, heads, groups):
    """modules/transformer_2d.py:216-399 (continuous input, conv projections) around the write branch of the
    hacked BasicTransformerBlock forward (modules/mutual_self_attention.py:127-130,145-174, feed-forward tail
    :270-283).  x (B,C,h,w), enc (B,1,cross) -> (output (B,C,h,w), bank (B,N,C) = norm2(h + attn1(norm1(h))))."""
    B, C, H, W = x.shape
    res = x
    h = group_norm(sd, p + ".norm", x, groups, 1e-6)
    h = conv(sd, p + ".proj_in", h, padding=0)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    tb = p + ".transformer_blocks.0"
    n = layer_norm(sd, tb + ".norm1", h)
    h = attention(sd, tb + ".attn1", n, n, heads) + h
    n = layer_norm(sd, tb + ".norm2", h)
    bank = n.clone()
    h = attention(sd, tb + ".attn2", n, enc, heads) + h
    h = feed_forward(sd, tb + ".ff", layer_norm(sd, tb + ".norm3", h)) + h
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    h = conv(sd, p + ".proj_out", h, padding=0)
    return h + res, bank


def refnet_forward(sd: Dict[str, Tensor], cfg, ref_latents: Tensor, taps: Optional[dict] = None):
    """The reference's ReferenceNet call (pipelines/v_express_pipeline.py:502-508):
    ``reference_net(ref_latents, timestep=0, encoder_hidden_states=zeros(1,1,cross))`` with the write-mode hooks
    installed -- modules/unet_2d_condition.py:877-1313 (forward), unet_2d_blocks.py (CrossAttnDownBlock2D :630-676,
    DownBlock2D :745-775, UNetMidBlock2DCrossAttn :470-507, UpBlock2D :1027-1073, CrossAttnUpBlock2D :890-961).

    ref_latents (1,4,h,w) -> (banks, out): ``banks`` is the list of the 16 (1,N,C) bank tensors in PAIRING order
    (``bank_order``: stable sort of the writer's dfs order by -dim, mutual_self_attention.py:349-351 -- the writer
    registers down_blocks, up_blocks, mid_block in the same order as the reader), ``out`` the (unused) UNet output."""
    B, cin, H, W = ref_latents.shape
    boc = cfg["block_out_channels"]
    heads, groups, eps = cfg["heads"], cfg["norm_num_groups"], cfg["norm_eps"]
    enc = torch.zeros(B, 1, cfg["cross_attention_dim"], dtype=ref_latents.dtype)
    bank_of: Dict[str, Tensor] = {}

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    t = torch.zeros(B, dtype=torch.long)                       # timestep=0 (unet_2d_condition.py:1010-1030)
    t_emb = timestep_embedding(t, boc[0]).to(ref_latents.dtype)
    emb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", t_emb)))

    def attn(p, x):
        x, bank_of[p] = transformer_2d_write(sd, p, x, enc, heads, groups)
        return x

    x = conv(sd, "conv_in", ref_latents)
    tap("conv_in", x)
    skips = [x]
    for i in range(4):
        p = f"down_blocks.{i}"
        for j in range(cfg["layers_per_block"]):
            x = resnet_block(sd, f"{p}.resnets.{j}", x, emb, groups, eps)
            if i < 3:
                x = attn(f"{p}.attentions.{j}", x)
            skips.append(x)
        if i < 3:
            x = conv(sd, f"{p}.downsamplers.0.conv", x, stride=2, padding=1)
            skips.append(x)
        tap(p, x)
    x = resnet_block(sd, "mid_block.resnets.0", x, emb, groups, eps)
    x = attn("mid_block.attentions.0", x)
    x = resnet_block(sd, "mid_block.resnets.1", x, emb, groups, eps)
    tap("mid_block", x)
    for i in range(4):
        p = f"up_blocks.{i}"
        for j in range(cfg["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"{p}.resnets.{j}", x, emb, groups, eps)
            if i > 0:
                x = attn(f"{p}.attentions.{j}", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv(sd, f"{p}.upsamplers.0.conv", x)
        tap(p, x)
    # the reference resets ``self.conv_norm_out = None`` after building it (unet_2d_condition.py:650), so its
    # state_dict has no conv_norm_out and forward (:1301-1304) applies conv_out directly
    x = conv(sd, "conv_out", x)
    return [bank_of[n] for n in bank_order(cfg)], x


def refnet_param_shapes(cfg) -> Dict[str, tuple]:
    """state_dict key -> shape of the reference's UNet2DConditionModel in the SD-1.5 configuration
    (modules/unet_2d_condition.py:69-660; 684 tensors, 859.5 M parameters at full width)."""
    boc = cfg["block_out_channels"]
    ted = boc[0] * 4
    cross = cfg["cross_attention_dim"]
    S: Dict[str, tuple] = {}

    def conv_(p, co, ci, k):
        S[p + ".weight"] = (co, ci, k, k)
        S[p + ".bias"] = (co,)

    def lin_(p, co, ci, bias=True):
        S[p + ".weight"] = (co, ci)
        if bias:
            S[p + ".bias"] = (co,)

    def norm_(p, c):
        S[p + ".weight"] = (c,)
        S[p + ".bias"] = (c,)

    def resnet_(p, ci, co):
        norm_(p + ".norm1", ci)
        conv_(p + ".conv1", co, ci, 3)
        lin_(p + ".time_emb_proj", co, ted)
        norm_(p + ".norm2", co)
        conv_(p + ".conv2", co, co, 3)
        if ci != co:
            conv_(p + ".conv_shortcut", co, ci, 1)

    def attn_(p, c, kv):
        lin_(p + ".to_q", c, c, False)
        lin_(p + ".to_k", c, kv, False)
        lin_(p + ".to_v", c, kv, False)
        lin_(p + ".to_out.0", c, c)

    def t2d_(p, c):
        norm_(p + ".norm", c)
        conv_(p + ".proj_in", c, c, 1)
        tb = p + ".transformer_blocks.0"
        norm_(tb + ".norm1", c)
        attn_(tb + ".attn1", c, c)
        norm_(tb + ".norm2", c)
        attn_(tb + ".attn2", c, cross)
        norm_(tb + ".norm3", c)
        lin_(tb + ".ff.net.0.proj", 8 * c, c)
        lin_(tb + ".ff.net.2", c, 4 * c)
        conv_(p + ".proj_out", c, c, 1)

    conv_("conv_in", boc[0], cfg["in_channels"], 3)
    lin_("time_embedding.linear_1", ted, boc[0])
    lin_("time_embedding.linear_2", ted, ted)
    out_c = boc[0]
    for i in range(4):
        in_c, out_c = out_c, boc[i]
        for j in range(cfg["layers_per_block"]):
            resnet_(f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
            if i < 3:
                t2d_(f"down_blocks.{i}.attentions.{j}", out_c)
        if i < 3:
            conv_(f"down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    c = boc[-1]
    resnet_("mid_block.resnets.0", c, c)
    t2d_("mid_block.attentions.0", c)
    resnet_("mid_block.resnets.1", c, c)
    rev = list(reversed(boc))
    out_c = rev[0]
    for i in range(4):
        prev_out = out_c
        out_c = rev[i]
        in_c = rev[min(i + 1, 3)]
        n = cfg["layers_per_block"] + 1
        for j in range(n):
            skip_c = in_c if j == n - 1 else out_c
            res_in = prev_out if j == 0 else out_c
            resnet_(f"up_blocks.{i}.resnets.{j}", res_in + skip_c, out_c)
            if i > 0:
                t2d_(f"up_blocks.{i}.attentions.{j}", out_c)
        if i < 3:
            conv_(f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    conv_("conv_out", cfg["out_channels"], boc[0], 3)     # no conv_norm_out: see refnet_forward
    return S


# --------------------------------------------------------------------------------------
# Conditioning prologue and post-processing (row f3)
# --------------------------------------------------------------------------------------

KPS_CFG = dict(conditioning_embedding_channels=320, conditioning_channels=3, block_out_channels=(16, 32, 96, 256))
AUDIO_PROJ_CFG = dict(dim=768, depth=4, dim_head=64, heads=12, num_queries=5, embedding_dim=768, output_dim=768,
                      ff_mult=4, max_seq_len=10)          # inference.py:116-126,192-201 with num_pad_audio_frames=2


def kps_guider_param_shapes(cfg=KPS_CFG) -> Dict[str, tuple]:
    """modules/v_kps_guider.py:10-33: conv_in, (same-width conv, stride-2 conv) per level, conv_out."""
    boc = cfg["block_out_channels"]
    S: Dict[str, tuple] = {}

    def conv_(p, co, ci):
        S[p + ".weight"] = (co, ci, 3, 3)
        S[p + ".bias"] = (co,)

    conv_("conv_in", boc[0], cfg["conditioning_channels"])
    for i in range(len(boc) - 1):
        conv_(f"blocks.{2 * i}", boc[i], boc[i])
        conv_(f"blocks.{2 * i + 1}", boc[i + 1], boc[i])
    conv_("conv_out", cfg["conditioning_embedding_channels"], boc[-1])
    return S


def kps_guider_forward(sd, cfg, images: Tensor) -> Tensor:
    """modules/v_kps_guider.py:35-45 on (b,3,t,H,W) keypoint images in [0,1]: every conv is an InflatedConv3d, i.e. a
    per-frame 2-D conv (modules/resnet.py:9-17); SiLU after every conv but the last.  -> (b,C0,t,H/8,W/8)."""
    b, c, t, H, W = images.shape
    x = images.permute(0, 2, 1, 3, 4).reshape(b * t, c, H, W)
    x = F.silu(conv(sd, "conv_in", x))
    for i in range(2 * (len(cfg["block_out_channels"]) - 1)):
        x = F.silu(conv(sd, f"blocks.{i}", x, stride=1 + (i & 1)))
    x = conv(sd, "conv_out", x)
    return x.view(b, t, x.shape[1], x.shape[2], x.shape[3]).permute(0, 2, 1, 3, 4).contiguous()


def audio_frame_windows(emb: Tensor, video_length: int, num_pad_audio_frames: int = 2) -> Tensor:
    """pipelines/v_express_pipeline.py:380-401: wav2vec2 states (1,T,d) -> linear interpolation (fp32) to 2*L steps,
    2*num_pad zero rows on both sides, and the (L, 2*(2*num_pad+1), d) sliding windows of the frames."""
    dt = emb.dtype
    x = F.interpolate(emb.float().permute(0, 2, 1), size=2 * video_length, mode="linear")[0].permute(1, 0).to(dt)
    pad = torch.zeros(2 * num_pad_audio_frames, x.shape[1], dtype=dt)
    x = torch.cat([pad, x, pad], 0)
    return torch.stack([x[2 * i:2 * (i + 2 * num_pad_audio_frames + 1)] for i in range(video_length)], 0)


def audio_projection_param_shapes(cfg=AUDIO_PROJ_CFG) -> Dict[str, tuple]:
    """modules/audio_projection.py:88-126 (num_latents_mean_pooled = 0)."""
    d, inner = cfg["dim"], cfg["dim_head"] * cfg["heads"]
    S: Dict[str, tuple] = {"pos_emb.weight": (cfg["max_seq_len"], cfg["embedding_dim"]), "latents": (1, cfg["num_queries"], d),
                           "proj_in.weight": (d, cfg["embedding_dim"]), "proj_in.bias": (d,),
                           "proj_out.weight": (cfg["output_dim"], d), "proj_out.bias": (cfg["output_dim"],),
                           "norm_out.weight": (cfg["output_dim"],), "norm_out.bias": (cfg["output_dim"],)}
    for i in range(cfg["depth"]):
        a, f = f"layers.{i}.0", f"layers.{i}.1"
        for n in ("norm1", "norm2"):
            S[f"{a}.{n}.weight"] = (d,)
            S[f"{a}.{n}.bias"] = (d,)
        S[a + ".to_q.weight"] = (inner, d)
        S[a + ".to_kv.weight"] = (2 * inner, d)
        S[a + ".to_out.weight"] = (d, inner)
        S[f + ".0.weight"] = (d,)
        S[f + ".0.bias"] = (d,)
        S[f + ".1.weight"] = (cfg["ff_mult"] * d, d)
        S[f + ".3.weight"] = (d, cfg["ff_mult"] * d)
    return S


def audio_projection_forward(sd, cfg, x: Tensor) -> Tensor:
    """modules/audio_projection.py:128-150 (perceiver resampler): x (L, 10, 768) -> (L, 5, 768).
    PerceiverAttention :44-75: q from the latents, k/v from [x | latents], both pre-normalised; softmax in fp32."""
    heads = cfg["heads"]
    n = x.shape[1]
    x = x + sd["pos_emb.weight"][:n]
    lat = sd["latents"].repeat(x.shape[0], 1, 1)
    x = _lin(sd, "proj_in", x)

    def split(t):
        return t.view(t.shape[0], t.shape[1], heads, -1).transpose(1, 2)

    for i in range(cfg["depth"]):
        a, f = f"layers.{i}.0", f"layers.{i}.1"
        xn = layer_norm(sd, a + ".norm1", x)
        ln = layer_norm(sd, a + ".norm2", lat)
        q = split(_lin(sd, a + ".to_q", ln))
        k, v = _lin(sd, a + ".to_kv", torch.cat([xn, ln], dim=-2)).chunk(2, dim=-1)
        k, v = split(k), split(v)
        sc = 1 / math.sqrt(math.sqrt(cfg["dim_head"]))
        w = torch.softmax(((q * sc) @ (k * sc).transpose(-2, -1)).float(), dim=-1).to(q.dtype)
        o = (w @ v).permute(0, 2, 1, 3).reshape(lat.shape[0], lat.shape[1], -1)
        lat = _lin(sd, a + ".to_out", o) + lat
        h = layer_norm(sd, f + ".0", lat)
        lat = _lin(sd, f + ".3", F.gelu(_lin(sd, f + ".1", h))) + lat
    return layer_norm(sd, "norm_out", _lin(sd, "proj_out", lat))


def median_filter_3d(video: Tensor, kernel_size: int = 3) -> Tensor:
    """pipelines/utils.py:46-63: (3,T,H,W) -> (3,T,H,W), median over the k x k x k neighbourhood (t,h,w) with reflect
    padding on all three axes.  k^3 is odd, so torch.median's lower-median rule coincides with the true median."""
    p = kernel_size // 2
    x = F.pad(video, (p, p, p, p, p, p), mode="reflect")
    _, T, H, W = video.shape
    views = [x[:, dt:dt + T, dy:dy + H, dx:dx + W] for dt in range(kernel_size) for dy in range(kernel_size)
             for dx in range(kernel_size)]
    return torch.stack(views, dim=-1).median(dim=-1)[0]


def video_to_uint8(video: Tensor) -> np.ndarray:
    """pipelines/utils.py:72-73: (3,T,H,W) in [0,1] -> (T,H,W,3) uint8 by truncation of v*255."""
    return (video.permute(1, 2, 3, 0) * 255).numpy().astype(np.uint8)


# --------------------------------------------------------------------------------------
# AutoencoderKL decoder (diffusers 0.29.2, sd-vae-ft-mse config)
# --------------------------------------------------------------------------------------

VAE_CFG = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4,
               norm_num_groups=32, out_channels=3)


def small_vae_cfg(width=(64, 64, 128, 128)):
    c = dict(VAE_CFG)
    c["block_out_channels"] = tuple(width)
    return c


def _vae_resnet(sd, p, x, groups):
    h = F.silu(group_norm(sd, p + ".norm1", x, groups, 1e-6))
    h = conv(sd, p + ".conv1", h)
    h = F.silu(group_norm(sd, p + ".norm2", h, groups, 1e-6))
    h = conv(sd, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in sd:
        x = conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def _vae_attn(sd, p, x, groups):
    B, C, H, W = x.shape
    res = x
    h = F.group_norm(x.view(B, C, H * W), groups, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], 1e-6)
    h = h.transpose(1, 2)
    q = _lin(sd, p + ".to_q", h)
    k = _lin(sd, p + ".to_k", h)
    v = _lin(sd, p + ".to_v", h)
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = _lin(sd, p + ".to_out.0", o)
    return o.transpose(1, 2).reshape(B, C, H, W) + res


def vae_decode(sd, cfg, z: Tensor) -> Tensor:
    """AutoencoderKL.decode(z).sample: post_quant_conv -> Decoder (B.6). z (n,4,h,w) -> (n,3,8h,8w)."""
    g = cfg["norm_num_groups"]
    nb = len(cfg["block_out_channels"])
    x = conv(sd, "post_quant_conv", z, padding=0)
    d = "decoder"
    x = conv(sd, d + ".conv_in", x)
    x = _vae_resnet(sd, d + ".mid_block.resnets.0", x, g)
    x = _vae_attn(sd, d + ".mid_block.attentions.0", x, g)
    x = _vae_resnet(sd, d + ".mid_block.resnets.1", x, g)
    for i in range(nb):
        for j in range(cfg["layers_per_block"] + 1):
            x = _vae_resnet(sd, f"{d}.up_blocks.{i}.resnets.{j}", x, g)
        if i < nb - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv(sd, f"{d}.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(group_norm(sd, d + ".conv_norm_out", x, g, 1e-6))
    return conv(sd, d + ".conv_out", x)


def vae_encode_mean(sd, cfg, x: Tensor) -> Tensor:
    """AutoencoderKL.encode(x).latent_dist.mean (diffusers 0.29.2 Encoder, sd-vae-ft-mse layout; the call the reference
    makes at pipelines/v_express_pipeline.py:346): conv_in -> DownEncoderBlock2D x len(boc) (resnets, then
    Downsample2D(padding=0): F.pad (0,1,0,1) + conv3x3 stride 2) -> UNetMidBlock2D -> GroupNorm -> SiLU -> conv_out ->
    quant_conv; mean = first latent_channels channels.  x (n,3,H,W) in [-1,1] -> (n,4,H/8,W/8)."""
    g = cfg["norm_num_groups"]
    boc = cfg["block_out_channels"]
    e = "encoder"
    h = conv(sd, e + ".conv_in", x)
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"]):
            h = _vae_resnet(sd, f"{e}.down_blocks.{i}.resnets.{j}", h, g)
        if i < len(boc) - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = conv(sd, f"{e}.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0)
    h = _vae_resnet(sd, e + ".mid_block.resnets.0", h, g)
    h = _vae_attn(sd, e + ".mid_block.attentions.0", h, g)
    h = _vae_resnet(sd, e + ".mid_block.resnets.1", h, g)
    h = F.silu(group_norm(sd, e + ".conv_norm_out", h, g, 1e-6))
    moments = conv(sd, "quant_conv", conv(sd, e + ".conv_out", h), padding=0)
    return moments[:, :cfg["latent_channels"]]


def vae_encoder_param_shapes(cfg) -> Dict[str, tuple]:
    """``encoder.*`` + ``quant_conv.*`` keys of diffusers' AutoencoderKL."""
    boc = cfg["block_out_channels"]
    S: Dict[str, tuple] = {}

    def conv_(p, co, ci, k):
        S[p + ".weight"] = (co, ci, k, k)
        S[p + ".bias"] = (co,)

    def norm_(p, c):
        S[p + ".weight"] = (c,)
        S[p + ".bias"] = (c,)

    def res_(p, ci, co):
        norm_(p + ".norm1", ci)
        conv_(p + ".conv1", co, ci, 3)
        norm_(p + ".norm2", co)
        conv_(p + ".conv2", co, co, 3)
        if ci != co:
            conv_(p + ".conv_shortcut", co, ci, 1)

    lc = cfg["latent_channels"]
    e = "encoder"
    conv_(e + ".conv_in", boc[0], 3, 3)
    out_c = boc[0]
    for i, ch in enumerate(boc):
        in_c, out_c = out_c, ch
        for j in range(cfg["layers_per_block"]):
            res_(f"{e}.down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
        if i < len(boc) - 1:
            conv_(f"{e}.down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    top = boc[-1]
    res_(e + ".mid_block.resnets.0", top, top)
    a = e + ".mid_block.attentions.0"
    norm_(a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        S[f"{a}.{n}.weight"] = (top, top)
        S[f"{a}.{n}.bias"] = (top,)
    res_(e + ".mid_block.resnets.1", top, top)
    norm_(e + ".conv_norm_out", top)
    conv_(e + ".conv_out", 2 * lc, top, 3)
    conv_("quant_conv", 2 * lc, 2 * lc, 1)
    return S


def decode_latents(vae_sd, vae_cfg, latents: Tensor) -> Tensor:
    """pipelines/v_express_pipeline.py:152-166: z/0.18215, per-frame decode, (x/2+.5).clamp, fp32."""
    f = latents.shape[2]
    z = (1 / 0.18215) * latents
    z = z.permute(0, 2, 1, 3, 4).reshape(-1, *z.shape[1:2], *z.shape[3:])
    frames = []
    for i in range(z.shape[0]):
        img = vae_decode(vae_sd, vae_cfg, z[i:i + 1])
        frames.append((img / 2 + 0.5).clamp(0, 1).float())
    v = torch.cat(frames)
    return v.view(-1, f, *v.shape[1:]).permute(0, 2, 1, 3, 4).contiguous()


# --------------------------------------------------------------------------------------
# Denoise loop (pipelines/v_express_pipeline.py:409-589, mean_overlap)
# --------------------------------------------------------------------------------------

def denoise(sd, cfg, latents: Tensor, kps_feature: Tensor, audio_embeddings: Tensor, banks,
            num_inference_steps: int, guidance_scale: float, context_frames: int, context_overlap: int,
            ref_w=1.0, audio_w=1.0, unet_fn=None) -> Tensor:
    """Steps x windows driver with overlap averaging + CFG + DDIM, same order of operations as the
    reference streaming loop (:527-572).  latents (1,4,L,h,w); kps (b,320,L,h,w); audio (b,L,5,768)."""
    sched = DDIM()
    sched.set_timesteps(num_inference_steps)
    L = latents.shape[2]
    do_cfg = guidance_scale > 1.0
    windows = context_windows(L, context_frames, context_overlap)
    nfc = torch.from_numpy(num_frame_context(windows, L))
    latents = latents.clone()
    if unet_fn is None:
        def unet_fn(x, t, enc, kps):
            return unet_forward(sd, cfg, x, t, enc, kps, banks, ref_w, audio_w)
    for t in sched.timesteps:
        counter = torch.zeros(L, dtype=torch.long)
        noise_preds = [None] * L
        for window in windows:
            kps = kps_feature[:, :, window]
            aud = audio_embeddings[:, window]
            aud = aud.reshape(-1, aud.shape[-2], aud.shape[-1])
            x = latents[:, :, window].repeat(2 if do_cfg else 1, 1, 1, 1, 1)
            noise_pred = unet_fn(x, t, aud, kps)
            if do_cfg:
                u, c = noise_pred.chunk(2)
                noise_pred = u + guidance_scale * (c - u)
            counter[window] += 1
            noise_pred = noise_pred / nfc[window][None, None, :, None, None]
            ids, preds = [], []
            for li, fi in enumerate(window):
                if noise_preds[fi] is None:
                    noise_preds[fi] = noise_pred[:, :, li].clone()
                else:
                    noise_preds[fi] += noise_pred[:, :, li]
                if counter[fi] == nfc[fi]:
                    ids.append(fi)
                    preds.append(noise_preds[fi])
                    noise_preds[fi] = None
            preds = torch.stack(preds, dim=2)
            stepped = sched.step(preds, t, latents[:, :, ids]).prev_sample
            # the reference writes `latents[:, :, ids] = stepped` (:572).  When a reflected window repeats a frame, `ids`
            # holds duplicates; torch leaves index_put_ with duplicates undefined (multi-threaded it is a race).  Defined
            # here as the sequential order -- the last write wins -- which is what the reference does on one thread.
            for k, fi in enumerate(ids):
                latents[:, :, fi] = stepped[:, :, k]
    return latents


# --------------------------------------------------------------------------------------
# Synthetic weights / inputs; deterministic by seed
# --------------------------------------------------------------------------------------

def unet_param_shapes(cfg) -> Dict[str, tuple]:
    """state_dict key -> shape in the reference layout (Appendix C)."""
    boc = cfg["block_out_channels"]
    ted = boc[0] * 4
    cross = cfg["cross_attention_dim"]
    S: Dict[str, tuple] = {}

    def conv_(p, co, ci, k):
        S[p + ".weight"] = (co, ci, k, k)
        S[p + ".bias"] = (co,)

    def lin_(p, co, ci, bias=True):
        S[p + ".weight"] = (co, ci)
        if bias:
            S[p + ".bias"] = (co,)

    def norm_(p, c):
        S[p + ".weight"] = (c,)
        S[p + ".bias"] = (c,)

    def resnet_(p, ci, co):
        norm_(p + ".norm1", ci)
        conv_(p + ".conv1", co, ci, 3)
        lin_(p + ".time_emb_proj", co, ted)
        norm_(p + ".norm2", co)
        conv_(p + ".conv2", co, co, 3)
        if ci != co:
            conv_(p + ".conv_shortcut", co, ci, 1)

    def attn_(p, c, kv, ):
        lin_(p + ".to_q", c, c, False)
        lin_(p + ".to_k", c, kv, False)
        lin_(p + ".to_v", c, kv, False)
        lin_(p + ".to_out.0", c, c)

    def ff_(p, c):
        lin_(p + ".net.0.proj", 8 * c, c)
        lin_(p + ".net.2", c, 4 * c)

    def t3d_(p, c):
        norm_(p + ".norm", c)
        conv_(p + ".proj_in", c, c, 1)
        tb = p + ".transformer_blocks.0"
        attn_(tb + ".attn1", c, c)
        norm_(tb + ".norm1", c)
        attn_(tb + ".attn1_5", c, c)
        norm_(tb + ".norm1_5", c)
        attn_(tb + ".attn2", c, cross)
        norm_(tb + ".norm2", c)
        ff_(tb + ".ff", c)
        norm_(tb + ".norm3", c)
        conv_(p + ".proj_out", c, c, 1)

    def mm_(p, c):
        p = p + ".temporal_transformer"
        norm_(p + ".norm", c)
        lin_(p + ".proj_in", c, c)
        tb = p + ".transformer_blocks.0"
        for i in range(2):
            attn_(tb + f".attention_blocks.{i}", c, c)
            S[tb + f".attention_blocks.{i}.pos_encoder.pe"] = (1, cfg["temporal_max_len"], c)
            norm_(tb + f".norms.{i}", c)
        ff_(tb + ".ff", c)
        norm_(tb + ".ff_norm", c)
        lin_(p + ".proj_out", c, c)

    conv_("conv_in", boc[0], cfg["in_channels"], 3)
    lin_("time_embedding.linear_1", ted, boc[0])
    lin_("time_embedding.linear_2", ted, ted)
    out_c = boc[0]
    for i in range(4):
        in_c, out_c = out_c, boc[i]
        for j in range(cfg["layers_per_block"]):
            resnet_(f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
            if i < 3:
                t3d_(f"down_blocks.{i}.attentions.{j}", out_c)
            mm_(f"down_blocks.{i}.motion_modules.{j}", out_c)
        if i < 3:
            conv_(f"down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    c = boc[-1]
    resnet_("mid_block.resnets.0", c, c)
    t3d_("mid_block.attentions.0", c)
    mm_("mid_block.motion_modules.0", c)
    resnet_("mid_block.resnets.1", c, c)
    rev = list(reversed(boc))
    out_c = rev[0]
    for i in range(4):
        prev_out = out_c
        out_c = rev[i]
        in_c = rev[min(i + 1, 3)]
        n = cfg["layers_per_block"] + 1
        for j in range(n):
            skip_c = in_c if j == n - 1 else out_c
            res_in = prev_out if j == 0 else out_c
            resnet_(f"up_blocks.{i}.resnets.{j}", res_in + skip_c, out_c)
            if i > 0:
                t3d_(f"up_blocks.{i}.attentions.{j}", out_c)
            mm_(f"up_blocks.{i}.motion_modules.{j}", out_c)
        if i < 3:
            conv_(f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    norm_("conv_norm_out", boc[0])
    conv_("conv_out", cfg["out_channels"], boc[0], 3)
    return S


def vae_param_shapes(cfg) -> Dict[str, tuple]:
    boc = cfg["block_out_channels"]
    S: Dict[str, tuple] = {}

    def conv_(p, co, ci, k):
        S[p + ".weight"] = (co, ci, k, k)
        S[p + ".bias"] = (co,)

    def norm_(p, c):
        S[p + ".weight"] = (c,)
        S[p + ".bias"] = (c,)

    def res_(p, ci, co):
        norm_(p + ".norm1", ci)
        conv_(p + ".conv1", co, ci, 3)
        norm_(p + ".norm2", co)
        conv_(p + ".conv2", co, co, 3)
        if ci != co:
            conv_(p + ".conv_shortcut", co, ci, 1)

    lc = cfg["latent_channels"]
    conv_("post_quant_conv", lc, lc, 1)
    d = "decoder"
    top = boc[-1]
    conv_(d + ".conv_in", top, lc, 3)
    res_(d + ".mid_block.resnets.0", top, top)
    a = d + ".mid_block.attentions.0"
    norm_(a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        S[f"{a}.{n}.weight"] = (top, top)
        S[f"{a}.{n}.bias"] = (top,)
    res_(d + ".mid_block.resnets.1", top, top)
    rev = list(reversed(boc))
    out_c = rev[0]
    for i in range(len(boc)):
        in_c, out_c = out_c, rev[i]
        for j in range(cfg["layers_per_block"] + 1):
            res_(f"{d}.up_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
        if i < len(boc) - 1:
            conv_(f"{d}.up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    norm_(d + ".conv_norm_out", boc[0])
    conv_(d + ".conv_out", cfg["out_channels"], boc[0], 3)
    return S


def synth_state_dict(shapes: Dict[str, tuple], seed: int) -> Dict[str, Tensor]:
    """weights randn/sqrt(fan_in), biases 0.02*randn, norm weights 1+0.02*randn,
    norm biases 0.02*randn, pos_encoder.pe = the deterministic table.  Per-key generators keyed by
    (seed, key) so any subset is reproducible independent of iteration order."""
    import zlib
    sd = {}
    for k, shp in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        if k.endswith("pos_encoder.pe"):
            sd[k] = positional_encoding(shp[2], shp[1])
        elif (len(k.split(".")) >= 2 and "norm" in k.split(".")[-2]) or (len(k.split(".")) > 2 and k.split(".")[-3] == "norms"):
            if k.endswith(".weight"):
                sd[k] = 1 + 0.02 * torch.randn(shp, generator=g)
            else:
                sd[k] = 0.02 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = int(np.prod(shp[1:]))
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(fan_in)
    return sd


def _ln_rows(x):
    return F.layer_norm(x, (x.shape[-1],))


def synth_inputs(cfg, L: int, h: int, w: int, do_cfg=True, seed=42):
    """Latents / kps features / audio tokens / banks per SURVEY.md 8(d)."""
    boc = cfg["block_out_channels"]
    b = 2 if do_cfg else 1
    g = lambda s: torch.Generator().manual_seed(s)
    latents = torch.randn(1, 4, L, h, w, generator=g(seed))
    kps = 0.1 * torch.randn(1, boc[0], L, h, w, generator=g(seed + 1))
    audio = _ln_rows(torch.randn(1, L, 5, cfg["cross_attention_dim"], generator=g(seed + 2)))
    if do_cfg:
        kps = torch.cat([torch.zeros_like(kps), kps], 0)
        audio = torch.cat([torch.zeros_like(audio), audio], 0)
    banks = []
    for i, name in enumerate(bank_order(cfg)):
        C = attention_block_dim(name, cfg)
        # token count of the level the block lives at
        if name.startswith("down_blocks."):
            s = int(name.split(".")[1])
        elif name.startswith("up_blocks."):
            s = 3 - int(name.split(".")[1])
        else:
            s = 3
        N = (h >> s) * (w >> s)
        v = _ln_rows(torch.randn(1, N, C, generator=g(seed + 3 + i)))
        if do_cfg:
            v = torch.cat([torch.zeros_like(v), v], 0)
        banks.append(v)
    return latents, kps, audio, banks
