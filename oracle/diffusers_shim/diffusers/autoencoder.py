import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))
from oracle import vx_oracle as O  # noqa: E402
from .configuration_utils import _Config  # noqa: E402


class _Out:
    def __init__(self, sample):
        self.sample = sample


class AutoencoderKL(torch.nn.Module):
    """Decoder-only AutoencoderKL over the restated functional decoder of the oracle."""

    def __init__(self, sd, cfg):
        super().__init__()
        self.cfg = cfg
        self.config = _Config(block_out_channels=cfg["block_out_channels"])
        self.params = torch.nn.ParameterDict({k.replace(".", "/"): torch.nn.Parameter(v, requires_grad=False)
                                              for k, v in sd.items()})

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def decode(self, z):
        sd = {k.replace("/", "."): v for k, v in self.params.items()}
        return _Out(O.vae_decode(sd, self.cfg, z))
