import torch


def maybe_allow_in_graph(cls):
    return cls


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)
