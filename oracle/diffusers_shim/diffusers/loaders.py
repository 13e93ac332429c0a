"""Name only: the ReferenceNet never loads LoRA / IP-adapter weights on this path."""


class UNet2DConditionLoadersMixin:
    pass
