"""diffusers.models.activations.get_activation restated (diffusers 0.29.2: name -> torch module)."""
from torch import nn

_ACT = {"swish": nn.SiLU, "silu": nn.SiLU, "mish": nn.Mish, "gelu": nn.GELU, "relu": nn.ReLU}


def get_activation(act_fn):
    return _ACT[act_fn.lower()]()
