from torch import nn


class AdaLayerNormSingle(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("PixArt-style norm: not used by the SD-1.5 ReferenceNet")
