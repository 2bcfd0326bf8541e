"""Post-load weight repacking: state_dict tensors (OIHW convs, [out,in] linears, reference key names) ->
the kernel-side layouts.  The state_dict itself is never modified (the dict is the contract).
"""
from __future__ import annotations

import torch

BF16 = torch.bfloat16
GEGLU_HALF = 128          # the GEGLU GEMM epilogue pairs column j with column j + 128 inside a 256-wide tile


def pack_conv3x3(w: torch.Tensor, dtype=BF16) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin] with K ordered (kh, kw, cin) = the implicit-GEMM K order (bf16 by default;
    dtype=None keeps the source precision for the check-mode (hi, lo) split)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    out = w.permute(0, 2, 3, 1).reshape(co, 9 * ci)
    return (out if dtype is None else out.to(dtype)).contiguous()


def pack_conv1x1(w: torch.Tensor, dtype=BF16) -> torch.Tensor:
    out = w.reshape(w.shape[0], w.shape[1])
    return (out if dtype is None else out.to(dtype)).contiguous()


def head_pitch(d: int) -> int:
    """Columns reserved per head in q/k/v buffers: d rounded up to the UMMA K step (16)."""
    return (d + 15) // 16 * 16


def pad_heads(w: torch.Tensor, heads: int) -> torch.Tensor:
    """[heads*d, in] -> [heads*pitch, in]: zero rows pad each head so padded q/k/v columns come out as exact
    zeros from the projection GEMM (they take part in the 16-wide tensor-core K steps)."""
    out_f, in_f = w.shape
    d = out_f // heads
    p = head_pitch(d)
    if p == d:
        return w
    wp = torch.zeros((heads, p, in_f), dtype=w.dtype, device=w.device)
    wp[:, :d] = w.reshape(heads, d, in_f)
    return wp.reshape(heads * p, in_f)


def interleave_geglu(w: torch.Tensor, b: torch.Tensor):
    """GEGLU proj [8C, C] (rows [0,4C) value, [4C,8C) gate; diffusers GEGLU.chunk(2)) -> per 256-row tile:
    128 value rows followed by their 128 gate rows."""
    two_n, k = w.shape
    n = two_n // 2
    assert n % GEGLU_HALF == 0, n
    val = w[:n].reshape(n // GEGLU_HALF, GEGLU_HALF, k)
    gate = w[n:].reshape(n // GEGLU_HALF, GEGLU_HALF, k)
    wi = torch.cat([val, gate], dim=1).reshape(two_n, k).contiguous()
    bv = b[:n].reshape(n // GEGLU_HALF, GEGLU_HALF)
    bg = b[n:].reshape(n // GEGLU_HALF, GEGLU_HALF)
    bi = torch.cat([bv, bg], dim=1).reshape(two_n).contiguous()
    return wi, bi


def rope_table(freqs: torch.Tensor, frames: int) -> torch.Tensor:
    """[F, rot_pairs, 2] (cos, sin) of position * freqs[i] (rotary_embedding_torch semantics)."""
    ang = torch.arange(frames, dtype=torch.float32, device=freqs.device)[:, None] * freqs.float()[None, :]
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()


def rel_pos_bias_table(emb: torch.Tensor, frames: int, num_buckets: int = 32, max_distance: int = 32) -> torch.Tensor:
    """RelativePositionBias.forward (base/models/attention.py:680-707) -> fp32 [heads, F, F].
    Integer bucket arithmetic done on the host once per (weights, F)."""
    import math
    pos = torch.arange(frames, device=emb.device)
    rel = pos[None, :] - pos[:, None]
    n = -rel
    half = num_buckets // 2
    ret = (n < 0).long() * half
    n = n.abs()
    max_exact = half // 2
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                         * (half - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, half - 1))
    bucket = ret + torch.where(n < max_exact, n, large)
    return emb.float()[bucket].permute(2, 0, 1).contiguous()


def pack_upsample_conv3x3(w: torch.Tensor, dtype=BF16) -> torch.Tensor:
    """Upsample3D's conv [Cout, Cin, 3, 3] -> [4*Cout, 4*Cin] for lavie_upsample_conv3x3_bf16: rows phase (py*2+px)-major,
    K ordered (a, b, cin).  A 3x3 conv on the nearest-2x upsampled map is, per output parity, a 2x2 conv on the source map:
    upsampled row 2y + py + kh - 1 is source row y + floor((py + kh - 1) / 2), so the taps that land on the same source
    pixel are summed (in fp32, before rounding)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    taps = {0: ([0], [1, 2]), 1: ([0, 1], [2])}          # parity -> (original taps of a = 0, of a = 1)
    wf = w.float()
    out = torch.empty((2, 2, co, 2, 2, ci), dtype=torch.float32, device=w.device)
    for py in range(2):
        for px in range(2):
            for a in range(2):
                for b in range(2):
                    acc = 0
                    for r in taps[py][a]:
                        for c in taps[px][b]:
                            acc = acc + wf[:, :, r, c]
                    out[py, px, :, a, b, :] = acc
    out = out.reshape(4 * co, 4 * ci)
    return (out if dtype is None else out.to(dtype)).contiguous()
