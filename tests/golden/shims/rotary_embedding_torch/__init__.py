"""Stand-in for `rotary_embedding_torch.RotaryEmbedding` (lucidrains), the subset
the reference calls (base/models/unet.py:185, base/models/attention.py:645-646).
TEST INFRASTRUCTURE ONLY."""
import torch
from torch import nn


def _rotate_half(x):
    x = x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)

    def forward(self, t):
        freqs = torch.einsum("..., f -> ... f", t.type(self.freqs.dtype), self.freqs)
        return freqs.repeat_interleave(2, dim=-1)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        seq_len = t.shape[seq_dim]
        freqs = self.forward(torch.arange(seq_len, device=t.device))
        rot = freqs.shape[-1]
        t_mid, t_right = t[..., :rot], t[..., rot:]
        t_mid = t_mid * freqs.cos() + _rotate_half(t_mid) * freqs.sin()
        return torch.cat((t_mid, t_right), dim=-1)
