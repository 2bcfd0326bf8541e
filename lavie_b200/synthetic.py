"""Deterministic random-init weights and inputs of the LaVie base architecture.

BASELINE.json asks for "random-init weights of the same architecture".  The
reference's own init depends on module construction order under a global seed and
cannot be reproduced without the reference, so every tensor is drawn here from a
generator keyed by (seed, crc32(key)): the same dict can be rebuilt anywhere
(container, GPU box) from the key/shape table alone.  Distributions follow
PyTorch's defaults (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for Linear/Conv weights and
biases, N(0,1) for nn.Embedding) with two deliberate differences that make parity
tests sensitive: norm affine parameters are perturbed away from (1, 0), and
``attn_temp.to_out.0.weight`` is NOT zero (the reference zero-inits it,
base/models/attention.py:475, which would hide the temporal attention).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

from .config import BASE_CONFIG, UNetConfig, param_spec


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def synthetic_state_dict(cfg: UNetConfig = BASE_CONFIG, seed: int = 0, dtype=torch.float32):
    sd = OrderedDict()
    spec = param_spec(cfg)
    for key, shape in spec.items():
        g = _gen(seed, key)
        if key.endswith("rotary_emb.freqs"):
            d = cfg.rotary_dim
            t = 1.0 / (10000.0 ** (torch.arange(0, d, 2)[: d // 2].float() / d))
        elif "relative_attention_bias" in key:
            t = torch.randn(shape, generator=g)
        elif ".norm" in key or key.startswith("conv_norm_out"):
            t = 0.1 * torch.randn(shape, generator=g)
            if key.endswith("weight"):
                t = t + 1.0
        else:
            if key.endswith("bias"):
                wshape = spec[key[: -len("bias")] + "weight"]
            else:
                wshape = shape
            fan_in = 1
            for s in wshape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        sd[key] = t.to(dtype)
    return sd


def synthetic_inputs(batch=2, frames=16, height=40, width=64, cfg: UNetConfig = BASE_CONFIG, seed: int = 0,
                     text_len: int = 77):
    """sample [B,C,F,H,W], integer timestep, encoder_hidden_states [B,77,ctx]."""
    g = _gen(seed, f"inputs:{batch}x{frames}x{height}x{width}x{text_len}")
    sample = torch.randn((batch, cfg.in_channels, frames, height, width), generator=g)
    text = torch.randn((batch, text_len, cfg.cross_attention_dim), generator=g)
    return sample, 500, text
