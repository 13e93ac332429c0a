import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields

SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
WEIGHTS_NAME = "diffusion_pytorch_model.bin"


class logging:  # noqa: N801
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


def is_accelerate_available():
    return False


USE_PEFT_BACKEND = False


def deprecate(*a, **k):
    pass


def scale_lora_layers(model, weight):
    pass


def unscale_lora_layers(model, weight=None):
    pass


def is_torch_version(op, version):
    import operator
    import torch
    from packaging import version as V
    ops = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "<": operator.lt, "<=": operator.le, "!=": operator.ne}
    return ops[op](V.parse(torch.__version__.split("+")[0]), V.parse(version))
