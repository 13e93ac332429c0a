"""Timesteps / TimestepEmbedding restated."""
import math

import torch
from torch import nn


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels, self.flip, self.shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", **_):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


class SinusoidalPositionalEmbedding(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError


def _unused(name):
    class _U(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(name + ": not used by the SD-1.5 ReferenceNet")
    _U.__name__ = name
    return _U


# names imported by modules/unet_2d_condition.py:18-33 and transformer_2d.py:7-10 but never constructed for SD-1.5
GaussianFourierProjection = _unused("GaussianFourierProjection")
ImageHintTimeEmbedding = _unused("ImageHintTimeEmbedding")
ImageProjection = _unused("ImageProjection")
ImageTimeEmbedding = _unused("ImageTimeEmbedding")
TextImageProjection = _unused("TextImageProjection")
TextImageTimeEmbedding = _unused("TextImageTimeEmbedding")
TextTimeEmbedding = _unused("TextTimeEmbedding")
PositionNet = _unused("PositionNet")
CaptionProjection = _unused("CaptionProjection")
ImagePositionalEmbeddings = _unused("ImagePositionalEmbeddings")
PatchEmbed = _unused("PatchEmbed")
