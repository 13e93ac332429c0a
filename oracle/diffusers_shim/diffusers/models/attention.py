"""FeedForward / GEGLU restated + stubs for unused names."""
import torch.nn.functional as F
from torch import nn

from .attention_processor import Attention  # noqa: F401


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", **_):
        super().__init__()
        assert activation_fn == "geglu"
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim)])

    def forward(self, x, *a, **k):
        for m in self.net:
            x = m(x)
        return x


class _Unused(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the hot path")


AdaLayerNorm = AdaLayerNormZero = GatedSelfAttentionDense = _Unused
