"""ResnetBlock2D / Downsample2D / Upsample2D of diffusers 0.29.2 restated for the configuration the SD-1.5
ReferenceNet uses (time_embedding_norm="default", swish, no up/down inside the resnet, conv down/up-samplers)."""
import torch
import torch.nn.functional as F
from torch import nn


class ResnetBlock2D(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512, groups=32,
                 groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", kernel=None, output_scale_factor=1.0, use_in_shortcut=None, up=False,
                 down=False, conv_shortcut_bias=True, conv_2d_out_channels=None):
        super().__init__()
        assert time_embedding_norm == "default" and not up and not down and non_linearity in ("swish", "silu")
        out_channels = in_channels if out_channels is None else out_channels
        self.output_scale_factor = output_scale_factor
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups_out or groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, conv_2d_out_channels or out_channels, 3, 1, 1)
        self.nonlinearity = nn.SiLU()
        use_in_shortcut = in_channels != (conv_2d_out_channels or out_channels) if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = (nn.Conv2d(in_channels, conv_2d_out_channels or out_channels, 1, 1, 0, bias=conv_shortcut_bias)
                              if use_in_shortcut else None)

    def forward(self, input_tensor, temb=None, scale=1.0):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        if self.time_emb_proj is not None and temb is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv", **_):
        super().__init__()
        assert use_conv
        self.padding = padding
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)

    def forward(self, hidden_states, scale=1.0):
        if self.padding == 0:
            hidden_states = F.pad(hidden_states, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(hidden_states)


class Upsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv", **_):
        super().__init__()
        assert use_conv and not use_conv_transpose
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, padding=1)

    def forward(self, hidden_states, output_size=None, scale=1.0):
        dtype = hidden_states.dtype
        if dtype == torch.bfloat16:
            hidden_states = hidden_states.to(torch.float32)
        if output_size is None:
            hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
        else:
            hidden_states = F.interpolate(hidden_states, size=output_size, mode="nearest")
        if dtype == torch.bfloat16:
            hidden_states = hidden_states.to(dtype)
        return self.conv(hidden_states)
