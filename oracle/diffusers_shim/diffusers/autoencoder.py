"""``AutoencoderKL`` restated as ``nn.Module``s in the structure of diffusers 0.29.2 (models/autoencoders/vae.py:
``Encoder`` / ``Decoder`` / ``UNetMidBlock2D`` / ``UpDecoderBlock2D`` / ``DownEncoderBlock2D``), built from the shim's
own leaf modules (``ResnetBlock2D``, ``Upsample2D``, ``Downsample2D``, ``Attention``) -- INDEPENDENTLY of the functional
decoder in ``oracle/vx_oracle.py``: the module tree is what defines the state_dict key layout, the forward is the
library's call order.  ``oracle/gen_golden.py`` runs the reference pipeline over THIS class; the oracle's functional
decoder and the product are then both held to it.  TEST INFRASTRUCTURE ONLY."""
import torch
from torch import nn

from .configuration_utils import _Config
from .models.attention_processor import Attention
from .models.resnet import Downsample2D, ResnetBlock2D, Upsample2D


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        return self.mean + self.std * torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype)

    def mode(self):
        return self.mean


class AutoencoderKLOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist


class UNetMidBlock2D(nn.Module):
    def __init__(self, in_channels, resnet_eps=1e-6, resnet_groups=32, attention_head_dim=None):
        super().__init__()
        attention_head_dim = attention_head_dim or in_channels
        mk = lambda: ResnetBlock2D(in_channels=in_channels, out_channels=in_channels, temb_channels=None, eps=resnet_eps,
                                   groups=resnet_groups)
        self.resnets = nn.ModuleList([mk(), mk()])
        self.attentions = nn.ModuleList([Attention(in_channels, heads=in_channels // attention_head_dim,
                                                   dim_head=attention_head_dim, rescale_output_factor=1.0, eps=resnet_eps,
                                                   norm_num_groups=resnet_groups, residual_connection=True, bias=True,
                                                   upcast_softmax=True)])

    def forward(self, hidden_states, temb=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, temb=temb)
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample, resnet_eps=1e-6, resnet_groups=32):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels,
                                                    out_channels=out_channels, temb_channels=None, eps=resnet_eps,
                                                    groups=resnet_groups) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None

    def forward(self, hidden_states, temb=None):
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb=temb)
        if self.upsamplers is not None:
            for up in self.upsamplers:
                hidden_states = up(hidden_states)
        return hidden_states


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample, resnet_eps=1e-6, resnet_groups=32):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels,
                                                    out_channels=out_channels, temb_channels=None, eps=resnet_eps,
                                                    groups=resnet_groups) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels, padding=0)]) \
            if add_downsample else None

    def forward(self, hidden_states):
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb=None)
        if self.downsamplers is not None:
            for down in self.downsamplers:
                hidden_states = down(hidden_states)
        return hidden_states


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[-1], 3, 1, 1)
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], resnet_groups=norm_num_groups)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(block_out_channels))
        out_ch = rev[0]
        for i in range(len(rev)):
            prev, out_ch = out_ch, rev[i]
            self.up_blocks.append(UpDecoderBlock2D(prev, out_ch, layers_per_block + 1, add_upsample=i != len(rev) - 1,
                                                   resnet_groups=norm_num_groups))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)

    def forward(self, sample):
        sample = self.conv_in(sample)
        sample = self.mid_block(sample)
        for up in self.up_blocks:
            sample = up(sample)
        return self.conv_out(self.conv_act(self.conv_norm_out(sample)))


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups, double_z=True):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, 1, 1)
        self.down_blocks = nn.ModuleList()
        out_ch = block_out_channels[0]
        for i, ch in enumerate(block_out_channels):
            prev, out_ch = out_ch, ch
            self.down_blocks.append(DownEncoderBlock2D(prev, out_ch, layers_per_block,
                                                       add_downsample=i != len(block_out_channels) - 1,
                                                       resnet_groups=norm_num_groups))
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], resnet_groups=norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[-1], 2 * out_channels if double_z else out_channels, 3, padding=1)

    def forward(self, sample):
        sample = self.conv_in(sample)
        for down in self.down_blocks:
            sample = down(sample)
        sample = self.mid_block(sample)
        return self.conv_out(self.conv_act(self.conv_norm_out(sample)))


class AutoencoderKL(nn.Module):
    """``AutoencoderKL(state_dict, cfg)``: cfg = dict(block_out_channels, layers_per_block, latent_channels,
    norm_num_groups).  Decoder always; the encoder half is built when the state_dict carries ``encoder.*`` keys."""

    def __init__(self, sd, cfg, in_channels=3, out_channels=3):
        super().__init__()
        boc = tuple(cfg["block_out_channels"])
        lc, g, lpb = cfg.get("latent_channels", 4), cfg.get("norm_num_groups", 32), cfg.get("layers_per_block", 2)
        self.config = _Config(block_out_channels=boc, scaling_factor=0.18215, latent_channels=lc)
        self.decoder = Decoder(lc, out_channels, boc, lpb, g)
        self.post_quant_conv = nn.Conv2d(lc, lc, 1)
        if any(k.startswith("encoder.") for k in sd):
            self.encoder = Encoder(in_channels, lc, boc, lpb, g)
            self.quant_conv = nn.Conv2d(2 * lc, 2 * lc, 1)
        missing, unexpected = self.load_state_dict(sd, strict=True), None
        for p in self.parameters():
            p.requires_grad_(False)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def encode(self, x):
        return AutoencoderKLOutput(DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))

    def decode(self, z):
        return DecoderOutput(self.decoder(self.post_quant_conv(z)))
