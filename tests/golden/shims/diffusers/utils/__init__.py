from collections import OrderedDict
from dataclasses import fields
from . import logging  # noqa: F401

WEIGHTS_NAME = "diffusion_pytorch_model.bin"


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


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.randn_tensor for the single-generator case the vendored scheduler uses."""
    import torch
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)
