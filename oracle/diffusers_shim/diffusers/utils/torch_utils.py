import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor: draw on the generator's device, then move."""
    gen_device = generator.device if generator is not None else (device or torch.device("cpu"))
    return torch.randn(shape, generator=generator, device=gen_device, dtype=dtype).to(device)


def apply_freeu(resolution_idx, hidden_states, res_hidden_states, **freeu_kwargs):
    raise NotImplementedError("FreeU is never enabled on this path")
