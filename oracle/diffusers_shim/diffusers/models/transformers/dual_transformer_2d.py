from torch import nn


class DualTransformer2DModel(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("dual cross-attention is not used by the SD-1.5 ReferenceNet")
