"""LoRACompatible{Conv,Linear} without LoRA: plain layers whose forward tolerates the `scale` argument."""
from torch import nn


class LoRACompatibleConv(nn.Conv2d):
    def forward(self, hidden_states, scale=1.0):
        return super().forward(hidden_states)


class LoRACompatibleLinear(nn.Linear):
    def forward(self, hidden_states, scale=1.0):
        return super().forward(hidden_states)
