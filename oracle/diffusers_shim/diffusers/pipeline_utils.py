import torch
from tqdm import tqdm


class DiffusionPipeline:
    def __init__(self):
        self._modules_reg = {}

    def register_modules(self, **kw):
        for k, v in kw.items():
            self._modules_reg[k] = v
            setattr(self, k, v)

    def _first_module(self):
        for v in self._modules_reg.values():
            if isinstance(v, torch.nn.Module):
                return v
        raise RuntimeError("no module")

    @property
    def device(self):
        return next(self._first_module().parameters()).device

    @property
    def dtype(self):
        return next(self._first_module().parameters()).dtype

    def progress_bar(self, iterable=None, total=None):
        return tqdm(iterable, total=total, disable=True)

    def to(self, *a, **k):
        for v in self._modules_reg.values():
            if isinstance(v, torch.nn.Module):
                v.to(*a, **k)
        return self
